// tcgen05 / TMEM implicit-GEMM 3x3 convolutions for sm_100a (NHWC fp32 in HBM, bf16x3 on the tensor cores).
//
// One CTA owns 128 output pixels (M = 128 accumulator rows = TMEM lanes: 16 groups of 8 consecutive
// columns) and NT <= 256 output channels (accumulator columns).  Data path per (tile, chunk of KC input
// channels):
//
//   HBM --TMA tiled load, fp32 halo box, hardware zero fill outside the image--> raw ring (128B / 64B swizzle)
//       --8 converter warps: fp32 -> (bf16 hi, bf16 lo)--> operand ring: 16-byte channel vectors, pixel-major
//         ("SWIZZLE_NONE core matrices": plane q = channels 8q..8q+7, 16 bytes per pixel)
//       --tcgen05.mma, one elected thread--> TMEM accumulators --4 epilogue warps--> alpha, bias, leaky-relu
//       --> swizzled staging tile --TMA tiled store (clipped at the tensor edge)--> HBM
//
// In the operand layout every filter tap is just a different START ADDRESS of the same staged tile (group
// stride fixed), so the nine taps cost nine descriptor pairs, not nine loads.  Each 16-channel K slice
// issues three MMAs (hi*hi + hi*lo + lo*hi); accumulation is fp32 in TMEM (the accumulator truncates:
// ~1e-5 relative at K = 2304, tools/tc_precision.py).  Weights are pre-split into the same layout by
// conv_tc_prep_kernel, contiguous per (output-channel tile, chunk), and streamed with 1-D bulk copies in
// groups of `tps` taps (or loaded once and kept when they fit).
// Warp roles: 0-3 and 16-19 epilogue (two groups splitting the channels of every chunk), 4-11 converters, 12 TMA halo
// loads, 13 weight copies, 14-15 MMA issue (14 also owns the TMEM allocation).  Rings: raw (TMA -> converters), A (converters -> MMA), B (weights),
// accumulators (per tile).
//
// MMA issue rate.  With 32 output channels one MMA is ~16 cycles of tensor work, less than a single warp
// needs to ISSUE it (measured: ~7 cycles per instruction of the issuing warp, profiles/).  Two remedies:
//  * "cat" mode: the weight planes are stored [q][hi | lo], so hi*[hi | lo] is ONE MMA of width 2*NT into
//    a double-width accumulator (the epilogue adds the halves) followed by lo*hi of width NT: two
//    instructions per K slice instead of three, same tensor work;
//  * two issuing warps that alternate TILES (each with its own accumulator buffer and its own half of the
//    operand / weight rings: stages w, w + 2, ... belong to issuing warp w, so every ring stays
//    single-producer single-consumer and the mbarrier parities never skip a phase).
//
// Tiles.  Images with >= 16 rows: 16 rows x 8 columns of one image.  Smaller images (8, 4, 2 rows: the
// 256-channel low-resolution blocks) interleave IMG = 16 / rows images row by row -- group g = row * IMG
// + slot -- which is exactly the order a TMA box over (c, w, n, h) produces.
//
// Forms (template FORM):
//   TC_C1  gather conv stride 1 (forward of conv2d; dgrad of stride 1 with flipped/transposed weights)
//   TC_C2  gather conv stride 2 (forward of the down-scaling conv; dgrad of conv2d_transpose); the
//          converters de-interleave the raw rows by column parity so that taps stay start addresses
//   TC_T2  transposed conv stride 2 (forward of conv2d_transpose; dgrad of the down-scaling conv):
//          four sub-pixel phases = four accumulators fed by 4/2/2/1 taps of one halo
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

enum { TC_C1 = 0, TC_C2 = 1, TC_T2 = 2 };

template <int FORM>
struct TcGeo {
  // pixels between consecutive 8-pixel groups of the staged tile
  static constexpr int GSTRIDE = (FORM == TC_C1) ? 10 : (FORM == TC_C2) ? 34 : 9;
  static constexpr int NACC = (FORM == TC_T2) ? 4 : 1;
  __host__ __device__ static constexpr int tap_acc(int t) {
    return (FORM == TC_T2) ? (((t / 3) == 1 ? 2 : 0) + ((t % 3) == 1 ? 1 : 0)) : 0;
  }
  // staged pixels for a tile of `rows` rows per image and `img` interleaved images
  __host__ __device__ static constexpr int pixels(int rows, int img) {
    return (FORM == TC_C1) ? (rows + 2) * img * 10 : (FORM == TC_C2) ? (rows + 1) * img * 34 : (rows + 1) * img * 9;
  }
  // raw (TMA box) rows / columns per image
  __host__ __device__ static constexpr int box_h(int rows) { return (FORM == TC_C1) ? rows + 2 : (FORM == TC_C2) ? 2 * rows + 1 : rows + 1; }
  __host__ __device__ static constexpr int box_w() { return (FORM == TC_C1) ? 10 : (FORM == TC_C2) ? 17 : 9; }
  // start offset (pixels) of tap t
  __host__ __device__ static constexpr int tap_off(int t, int img) {
    const int kh = t / 3, kw = t % 3;
    if (FORM == TC_C1) return kh * img * 10 + kw;
    if (FORM == TC_C2) return (kh == 0 ? 0 : kh == 1 ? 17 : img * 34) + (kw & 1) * 9 + (kw >> 1);
    return ((kh == 2) ? 0 : 1) * img * 9 + ((kw == 2) ? 0 : 1);
  }
};

struct TcParams {
  const __nv_bfloat16* wprep;     // [ndim/nt][kdim/KC][9][KC/8][2][nt][8]
  const float* bias;              // [ndim] or null
  float* y;                       // [n, h_out, w_out, ndim]
  int n_img, h_in, w_in, h_out, w_out, kdim, ndim;
  float alpha;
  int act;
  int rows, img;                  // tile: `rows` rows of each of `img` interleaved images, 8 columns
  int pix;                        // staged pixels per tile (operand ring)
  int rpix;                       // raw pixels per tile (TMA box)
  int tiles_h, tiles_w, ntiles;   // ntiles = image groups * tiles_h * tiles_w
  int nt, n_tiles;                // output-channel tile (grid.y)
  int tap_off[9];
  int sa, sb, ds;                 // ring depths: operand stages, weight stages, raw slots
  int cat;                        // hi x [hi | lo] as one double-width MMA (accumulator = 2 * nt columns)
  int nw;                         // MMA-issuing warps (1 or 2)
  int b_resident;                 // the weights of this CTA's channel tile are loaded once and kept
  int nbuf;                       // accumulator buffers (1 or 2)
  int tmem_cols;                  // power of two >= nbuf * NACC * nt * (1 + cat)
  uint32_t raw_slot_bytes;        // 1024-byte aligned
  int ksplit, cps;                // split-K: grid.z CTAs per (tile, channel tile), each `cps` channel chunks; partial
                                  // sums meet in y through TMA reduce-add (y pre-zeroed, activation applied afterwards)
  // fused epilogues (TcEpi).  MASK: y = lrelu'(aux) * alpha * conv (aux: a tensor shaped like y; its tiles arrive
  // through `tmaux` in a ring of `aux_k` 16 KB slots filled by the otherwise idle second issue warp).  PNF: y =
  // pixel_norm(lrelu(alpha * conv + bias)) over ALL output channels (needs nt == ndim, ksplit == 1),
  // rvec[pixel] = 1 / sqrt(mean_c(a^2) + eps).
  int epi, aux_k;
  float eps;
  float* rvec;                    // [n, h_out, w_out]
#ifdef GS_TC_PROF
  unsigned long long* prof;       // [role][wait0, wait1, wait2, total] cycles
#endif
};

enum TcEpi { TC_EPI_PLAIN = 0, TC_EPI_MASK = 1, TC_EPI_PNF = 2 };

// Output tensor maps: one per accumulator (sub-pixel phase of the transposed form; a strided view of y)
struct TcOutMaps {
  CUtensorMap m[4];
};

constexpr int TC_THREADS = 640;         // 16 warps of conv_tc's roles + a second epilogue warpgroup (warps 16-19)
constexpr int TC_MMA_WARP0 = 14, TC_MMA_WARPS = 2;
constexpr int TC_CONV_WARPS = 8;
constexpr int TC_MAX_STAGES = 4;       // raw and operand rings
constexpr int TC_MAX_BSTAGES = 24;     // weight ring (all (chunk, tap group) blocks when resident)
constexpr int TC_MAX_AUX = 4;          // aux ring of the MASK epilogue

// Weight pre-pass: W_eff[tap][k][n] (k = contraction channel, n = output channel) -> bf16 split blocks
// [n / nt][k / KC][tap][q][split][n % nt][e], value index k = KC*kc + 8*q + e.  w_is_kn: weight memory is
// [tap][k][n] (else [tap][n][k]); flip: use tap 8 - t (180-degree rotation).
template <int KC>
__device__ __forceinline__ void conv_tc_prep_elem(size_t i, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim,
                                                  int ndim, int nt, int w_is_kn, int flip) {
  constexpr int Q = KC / 8;
  const int nchunks = kdim / KC;
  int e = (int)(i % 8);
  size_t r = i / 8;
  int nl = (int)(r % nt);
  r /= nt;
  int q = (int)(r % Q);
  r /= Q;
  int tap = (int)(r % 9);
  r /= 9;
  int kc = (int)(r % nchunks);
  int ntile = (int)(r / nchunks);
  int k = kc * KC + q * 8 + e, n = ntile * nt + nl;
  int st = flip ? 8 - tap : tap;
  float v = w_is_kn ? w[((size_t)st * kdim + k) * ndim + n] : w[((size_t)st * ndim + n) * kdim + k];
  __nv_bfloat16 hi, lo;
  tc::split_bf16(v, hi, lo);
  const size_t plane = (size_t)nt * 8;                                   // elements of one [nt][8] plane
  const size_t blk = (((size_t)ntile * nchunks + kc) * 9 + tap) * (2 * Q);   // first plane of this (tile, chunk, tap)
  const size_t inner = (size_t)nl * 8 + e;
  out[(blk + 2 * q) * plane + inner] = hi;
  out[(blk + 2 * q + 1) * plane + inner] = lo;
}

// The same for the 8 consecutive contraction channels of group g = i / 8 (one 16-byte store per split term).
template <int KC>
__device__ __forceinline__ void conv_tc_prep_group(size_t g, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim,
                                                   int ndim, int nt, int w_is_kn, int flip) {
  constexpr int Q = KC / 8;
  const int nchunks = kdim / KC;
  size_t r = g;
  int nl = (int)(r % nt);
  r /= nt;
  int q = (int)(r % Q);
  r /= Q;
  int tap = (int)(r % 9);
  r /= 9;
  int kc = (int)(r % nchunks);
  int ntile = (int)(r / nchunks);
  const int k0 = kc * KC + q * 8, n = ntile * nt + nl;
  const int st = flip ? 8 - tap : tap;
  float v[8];
  if (w_is_kn) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = w[((size_t)st * kdim + k0 + e) * ndim + n];
  } else {
    const float4* src = reinterpret_cast<const float4*>(w + ((size_t)st * ndim + n) * kdim + k0);
    const float4 a = src[0], b = src[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  uint4 h4, l4;
  tc::split2_bf16(v[0], v[1], h4.x, l4.x);
  tc::split2_bf16(v[2], v[3], h4.y, l4.y);
  tc::split2_bf16(v[4], v[5], h4.z, l4.z);
  tc::split2_bf16(v[6], v[7], h4.w, l4.w);
  const size_t plane = (size_t)nt * 8;
  const size_t blk = (((size_t)ntile * nchunks + kc) * 9 + tap) * (2 * Q);
  *reinterpret_cast<uint4*>(out + (blk + 2 * q) * plane + (size_t)nl * 8) = h4;
  *reinterpret_cast<uint4*>(out + (blk + 2 * q + 1) * plane + (size_t)nl * 8) = l4;
}

template <int KC>
__global__ void conv_tc_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim, int ndim,
                                    int nt, int w_is_kn, int flip) {
  const size_t total = (size_t)9 * kdim * ndim;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    conv_tc_prep_elem<KC>(i, w, out, kdim, ndim, nt, w_is_kn, flip);
}

template <int FORM, int KC, int TPS, int CAT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ TcOutMaps tmy,
                                                                const __grid_constant__ TcOutMaps tmaux, const TcParams p) {
  using G = TcGeo<FORM>;
  constexpr int Q = KC / 8;               // 16-byte channel planes per split term
  extern __shared__ unsigned char tc_smem_raw[];
  __shared__ uint64_t raw_full[TC_MAX_STAGES], raw_empty[TC_MAX_STAGES], a_full[TC_MAX_STAGES], a_empty[TC_MAX_STAGES];
  __shared__ uint64_t b_full[TC_MAX_BSTAGES], b_empty[TC_MAX_BSTAGES];
  __shared__ uint64_t acc_full[2], acc_empty[2], aux_full[TC_MAX_AUX], aux_empty[TC_MAX_AUX];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[256];
  __shared__ float ss_x[2][128];          // PNF epilogue: per-pixel partial sums of squares of the two channel halves

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t plane_a = (uint32_t)p.pix * 16u;
  const uint32_t plane_b = (uint32_t)p.nt * 16u;
  const uint32_t a_stage_bytes = (uint32_t)(2 * Q) * plane_a;   // [split][q] planes
  const uint32_t b_tap_bytes = (uint32_t)(2 * Q) * plane_b;
  const uint32_t b_stage_bytes = (uint32_t)TPS * b_tap_bytes;
  // the swizzled TMA destination needs 1024-byte alignment
  unsigned char* tc_smem = tc_smem_raw + ((1024u - (tc::smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char* out_smem = tc_smem;                        // 2 x [128 pixels][32 channels] fp32, 128B-swizzled
  unsigned char* aux_smem = tc_smem + 2 * 16384;            // aux_k x [128 pixels][32 channels] fp32, 128B-swizzled
  unsigned char* raw_smem = aux_smem + (size_t)p.aux_k * 16384;
  unsigned char* a_smem = raw_smem + (size_t)p.ds * p.raw_slot_bytes;
  unsigned char* b_smem = a_smem + (size_t)p.sa * a_stage_bytes;
  uint16_t* dst_tab = reinterpret_cast<uint16_t*>(b_smem + (size_t)p.sb * b_stage_bytes);   // TC_C2 only
  const int nchunks_all = p.kdim / KC;
  const int kc_begin = blockIdx.z * p.cps;                     // this CTA's share of the contraction (split-K)
  const int kc_end = min(nchunks_all, kc_begin + p.cps);
  const int nchunks = kc_end - kc_begin;
  const int n0 = blockIdx.y * p.nt;       // first output channel of this CTA

  if (tid == 0) {
    for (int s = 0; s < p.ds; ++s) { tc::mbar_init(&raw_full[s], 1); tc::mbar_init(&raw_empty[s], TC_CONV_WARPS * 32); }
    for (int s = 0; s < p.sa; ++s) { tc::mbar_init(&a_full[s], TC_CONV_WARPS * 32); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.sb; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 256); }
    for (int s = 0; s < TC_MAX_AUX; ++s) { tc::mbar_init(&aux_full[s], 1); tc::mbar_init(&aux_empty[s], 256); }
    tc::mbar_fence_init();
  }
  for (int c = tid; c < p.nt; c += TC_THREADS) bias_s[c] = (p.bias && blockIdx.z == 0) ? p.bias[blockIdx.y * p.nt + c] : 0.0f;
  if (FORM == TC_C2) {
    // raw pixel (row hr, image slot, column hc) -> staged position: rows in pairs, columns split by parity
    for (int ps = tid; ps < p.rpix; ps += TC_THREADS) {
      const int hc = ps % 17, u = ps / 17;
      const int slot = u % p.img, hr = u / p.img;
      dst_tab[ps] = (uint16_t)(((hr >> 1) * p.img + slot) * 34 + (hr & 1) * 17 + (hc & 1) * 9 + (hc >> 1));
    }
  }
  if (warp == TC_MMA_WARP0) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp == 12 && lane == 0) {
    tc::prefetch_tmap(&tmx);
    for (int a = 0; a < G::NACC; ++a) tc::prefetch_tmap(&tmy.m[a]);
    if (p.epi == TC_EPI_MASK)
      for (int a = 0; a < G::NACC; ++a) tc::prefetch_tmap(&tmaux.m[a]);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int acc_w = CAT ? 2 * p.nt : p.nt;      // columns of one accumulator
  const int acc_cols = G::NACC * acc_w;           // columns per accumulator buffer

  if (warp >= 4 && warp < 12) {
    // ============================== fp32 -> bf16 hi/lo split ===========================================
    // item = (raw pixel, 8-channel group q): 32 bytes in, 16 + 16 bytes out.  A quarter warp reads 8
    // consecutive pixels of one q: the TMA swizzle makes that conflict-free, the stores are contiguous.
    const int cw = warp - 4;
    const int q = cw % Q;
    constexpr int PSTEP = 32 * (TC_CONV_WARPS / Q);
    const int p_first = (cw / Q) * 32 + lane;
    // operand ring: issuing warp w owns stages w, w + nw, ... (a private single-producer single-consumer ring)
    const int a_depth = p.sa / p.nw;
    int apos0 = 0, apos1 = 0, rs = 0, seq = 0;
    uint32_t aph0 = 0, aph1 = 0, rph = 0;
    TC_PROF_DECL
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++seq) {
      const bool w1 = (p.nw == 2) && (seq & 1);
      for (int kc = 0; kc < nchunks; ++kc) {
        const int stage = w1 ? 1 + 2 * apos1 : p.nw * apos0;
        const uint32_t aph = w1 ? aph1 : aph0;
        TC_WAIT(&raw_full[rs], rph, 0);
        TC_WAIT(&a_empty[stage], aph ^ 1u, 1);
        const unsigned char* raw = raw_smem + (size_t)rs * p.raw_slot_bytes;
        unsigned char* st = a_smem + (size_t)stage * a_stage_bytes;
        unsigned char* st_hi = st + (size_t)q * plane_a;
        unsigned char* st_lo = st + (size_t)(Q + q) * plane_a;
        // two items per pass: both loads are issued before either conversion
        for (int ps = p_first; ps < p.rpix; ps += 2 * PSTEP) {
          const int ps1 = ps + PSTEP;
          const bool two = ps1 < p.rpix;
          const int sw0 = (KC == 32) ? (ps & 7) : ((ps >> 1) & 3);
          const int sw1 = (KC == 32) ? (ps1 & 7) : ((ps1 >> 1) & 3);
          const unsigned char* row0 = raw + (size_t)ps * (KC * 4);
          const unsigned char* row1 = raw + (size_t)ps1 * (KC * 4);
          const float4 v0 = *reinterpret_cast<const float4*>(row0 + (((2 * q) ^ sw0) << 4));
          const float4 v1 = *reinterpret_cast<const float4*>(row0 + (((2 * q + 1) ^ sw0) << 4));
          float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
          if (two) {
            u0 = *reinterpret_cast<const float4*>(row1 + (((2 * q) ^ sw1) << 4));
            u1 = *reinterpret_cast<const float4*>(row1 + (((2 * q + 1) ^ sw1) << 4));
          }
          uint4 h4, l4, g4, k4;
          tc::split2_bf16(v0.x, v0.y, h4.x, l4.x);
          tc::split2_bf16(v0.z, v0.w, h4.y, l4.y);
          tc::split2_bf16(v1.x, v1.y, h4.z, l4.z);
          tc::split2_bf16(v1.z, v1.w, h4.w, l4.w);
          tc::split2_bf16(u0.x, u0.y, g4.x, k4.x);
          tc::split2_bf16(u0.z, u0.w, g4.y, k4.y);
          tc::split2_bf16(u1.x, u1.y, g4.z, k4.z);
          tc::split2_bf16(u1.z, u1.w, g4.w, k4.w);
          const int d0 = (FORM == TC_C2) ? (int)dst_tab[ps] : ps;
          *reinterpret_cast<uint4*>(st_hi + (size_t)d0 * 16) = h4;
          *reinterpret_cast<uint4*>(st_lo + (size_t)d0 * 16) = l4;
          if (two) {
            const int d1 = (FORM == TC_C2) ? (int)dst_tab[ps1] : ps1;
            *reinterpret_cast<uint4*>(st_hi + (size_t)d1 * 16) = g4;
            *reinterpret_cast<uint4*>(st_lo + (size_t)d1 * 16) = k4;
          }
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&a_full[stage]);
        tc::mbar_arrive(&raw_empty[rs]);
        if (w1) { if (++apos1 == a_depth) { apos1 = 0; aph1 ^= 1u; } }
        else { if (++apos0 == a_depth) { apos0 = 0; aph0 ^= 1u; } }
        if (++rs == p.ds) { rs = 0; rph ^= 1u; }
      }
    }
    TC_PROF_FLUSH(p.prof, 16, warp == 4 && lane == 0);
  } else if (warp == 12) {
    // ============================== halo tiles: one TMA box per (tile, chunk) ===========================
    if (lane == 0) {
      int rs = 0;
      uint32_t rph = 0;
      const uint32_t box_bytes = (uint32_t)p.rpix * (KC * 4u);
      TC_PROF_DECL
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int t = tile;
        const int tw_ = t % p.tiles_w;
        t /= p.tiles_w;
        const int th_ = t % p.tiles_h;
        const int img0 = (t / p.tiles_h) * p.img;
        const int w0 = (FORM == TC_C2) ? tw_ * 16 : tw_ * 8 - 1;
        const int h0 = (FORM == TC_C2) ? th_ * 32 : th_ * 16 - 1;
        for (int kc = 0; kc < nchunks; ++kc) {
          TC_WAIT(&raw_empty[rs], rph ^ 1u, 0);
          tc::mbar_arrive_expect_tx(&raw_full[rs], box_bytes);
          tc::tma_load_4d(raw_smem + (size_t)rs * p.raw_slot_bytes, &tmx, (kc_begin + kc) * KC, w0, img0, h0, &raw_full[rs]);
          if (++rs == p.ds) { rs = 0; rph ^= 1u; }
        }
      }
      TC_PROF_FLUSH(p.prof, 17, true);
    }
  } else if (warp == 13) {
    // ============================== weight blocks: one bulk copy per (chunk, tap group) =================
    if (lane == 0) {
      constexpr int groups = 9 / TPS;
      const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wprep) +
                                  ((size_t)blockIdx.y * nchunks_all + kc_begin) * 9 * b_tap_bytes;
      if (p.b_resident) {
        // one pass fills every stage; both issuing warps read the same copy
        for (int i = 0; i < nchunks * groups; ++i) {
          tc::mbar_arrive_expect_tx(&b_full[i], b_stage_bytes);
          tc::bulk_g2s(b_smem + (size_t)i * b_stage_bytes, wsrc + (size_t)i * b_stage_bytes, b_stage_bytes, &b_full[i]);
        }
      } else {
        // issuing warp w owns stages w, w + nw, ...
        const int b_depth = p.sb / p.nw;
        int bpos0 = 0, bpos1 = 0, seq = 0;
        uint32_t bph0 = 0, bph1 = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++seq) {
          const bool w1 = (p.nw == 2) && (seq & 1);
          for (int i = 0; i < nchunks * groups; ++i) {
            const int stage = w1 ? 1 + 2 * bpos1 : p.nw * bpos0;
            tc::mbar_wait(&b_empty[stage], (w1 ? bph1 : bph0) ^ 1u);
            tc::mbar_arrive_expect_tx(&b_full[stage], b_stage_bytes);
            tc::bulk_g2s(b_smem + (size_t)stage * b_stage_bytes, wsrc + (size_t)i * b_stage_bytes, b_stage_bytes, &b_full[stage]);
            if (w1) { if (++bpos1 == b_depth) { bpos1 = 0; bph1 ^= 1u; } }
            else { if (++bpos0 == b_depth) { bpos0 = 0; bph0 ^= 1u; } }
          }
        }
      }
    }
  } else if (warp >= TC_MMA_WARP0 && warp < TC_MMA_WARP0 + TC_MMA_WARPS) {
    // ============================== MMA issue ==========================================================
    // Descriptors are a per-stage base (start address in 16-byte units in the low word) plus tap / K-slice
    // / split offsets.  The WHOLE warp runs this control flow (so the compiler keeps descriptors in uniform
    // registers); one elected lane issues, once per weight stage.  Warp `mw` owns tiles mw, mw + nw, ...
    const int mw = warp - TC_MMA_WARP0;
    if (mw < p.nw) {
      constexpr int GROUPS = 9 / TPS;
      const uint32_t idesc_w = tc::idesc_bf16_f32(acc_w, 0, 0);      // hi x [hi | lo] in cat mode, else = idesc_n
      const uint32_t idesc_n = tc::idesc_bf16_f32(p.nt, 0, 0);
      const uint64_t a_desc0 = tc::smem_desc(tc::smem_u32(a_smem), plane_a, G::GSTRIDE * 16u);
      const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(b_smem), 2u * plane_b, 128u);
      const uint32_t a_stage16 = a_stage_bytes >> 4, b_stage16 = b_stage_bytes >> 4, b_tap16 = b_tap_bytes >> 4;
      const uint32_t plane_a16 = plane_a >> 4, plane_b16 = plane_b >> 4;
      const uint32_t a_lo16 = (uint32_t)Q * plane_a16;
      // private sub-rings: stages mw, mw + nw, ... of the operand ring (and of the weight ring when streamed)
      const int a_step = p.nw, b_step = p.b_resident ? 1 : p.nw;
      const int a_first = mw, b_first = p.b_resident ? 0 : mw;
      int sa = a_first, sb = b_first;
      uint32_t pa = 0, pb = 0;
      bool b_ready = false;                            // resident weights: waited for once
      int seq = mw;
      TC_PROF_DECL
      for (int tile = blockIdx.x + mw * gridDim.x; tile < p.ntiles; tile += p.nw * gridDim.x, seq += p.nw) {
        const int ab = seq % p.nbuf;
        const uint32_t pacc = (uint32_t)(seq / p.nbuf) & 1u;
        TC_WAIT(&acc_empty[ab], pacc ^ 1u, 0);
        tc::tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(ab * acc_cols);
        for (int kc = 0; kc < nchunks; ++kc) {
          TC_WAIT(&a_full[sa], pa, 1);
          tc::tc_fence_after();
          const uint64_t a_base = a_desc0 + (uint64_t)((uint32_t)sa * a_stage16);
          const uint32_t acc_rest = (kc > 0) ? 1u : 0u;
          const bool last_chunk = (kc == nchunks - 1);
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) {
            if (!b_ready) {
              TC_WAIT(&b_full[sb], pb, 2);
              tc::tc_fence_after();
            }
            const uint64_t b_base = b_desc0 + (uint64_t)((uint32_t)sb * b_stage16);
            if (tc::elect_one()) {
#pragma unroll
              for (int tt = 0; tt < TPS; ++tt) {
                const int tap = g * TPS + tt;
                const uint32_t d = d0 + (uint32_t)(G::tap_acc(tap) * acc_w);
                const uint64_t a_tap = a_base + (uint64_t)(uint32_t)p.tap_off[tap];
                const uint64_t b_tap = b_base + (uint64_t)((uint32_t)tt * b_tap16);
                // the first tap that touches an accumulator overwrites it on the first channel chunk
                const bool opens = (G::NACC == 1) ? (tap == 0) : (tap == 0 || tap == 1 || tap == 3 || tap == 4);
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  const uint64_t a_hi = a_tap + (uint64_t)((uint32_t)(2 * ks) * plane_a16);
                  const uint64_t b_hi = b_tap + (uint64_t)((uint32_t)(4 * ks) * plane_b16);
                  const uint32_t accum = (opens && ks == 0) ? acc_rest : 1u;
                  if (CAT) {
                    tc::mma_bf16(d, a_hi, b_hi, idesc_w, accum);
                    tc::mma_bf16(d, a_hi + a_lo16, b_hi, idesc_n, 1u);
                  } else {
                    tc::mma_bf16(d, a_hi, b_hi, idesc_n, accum);
                    tc::mma_bf16(d, a_hi, b_hi + plane_b16, idesc_n, 1u);
                    tc::mma_bf16(d, a_hi + a_lo16, b_hi, idesc_n, 1u);
                  }
                }
              }
              if (!p.b_resident) tc::mma_commit(&b_empty[sb]);
              if (g == GROUPS - 1) {
                tc::mma_commit(&a_empty[sa]);
                if (last_chunk) tc::mma_commit(&acc_full[ab]);
              }
            }
            __syncwarp();
            sb += b_step;
            if (sb >= p.sb) { sb = b_first; pb ^= 1u; }
          }
          sa += a_step;
          if (sa >= p.sa) { sa = a_first; pa ^= 1u; }
        }
        if (p.b_resident) b_ready = true;
      }
      TC_PROF_FLUSH(p.prof, 18, mw == 0 && lane == 0);
    } else if (p.epi == TC_EPI_MASK && lane == 0) {
      // ============================== aux tiles of the MASK epilogue (this warp issues no MMAs) ===========
      // one TMA box per (tile, accumulator, 32-channel chunk), in the order the epilogue consumes them
      int slot = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int t = tile;
        const int tw_ = t % p.tiles_w;
        t /= p.tiles_w;
        const int th_ = t % p.tiles_h;
        const int img0 = (t / p.tiles_h) * p.img;
        for (int a = 0; a < G::NACC; ++a)
          for (int c0 = 0; c0 < p.nt; c0 += 32) {
            tc::mbar_wait(&aux_empty[slot], ph ^ 1u);
            tc::mbar_arrive_expect_tx(&aux_full[slot], 16384u);
            tc::tma_load_4d(aux_smem + (size_t)slot * 16384, &tmaux.m[a], n0 + c0, tw_ * 8, img0, th_ * p.rows, &aux_full[slot]);
            if (++slot == p.aux_k) { slot = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp < 4 || warp >= 16) {
    // ============================== epilogue: TMEM -> alpha, bias, leaky-relu -> staging -> TMA store =====
    // Two warpgroups (warps 0-3 and 16-19) work on every unit (accumulator a, 32-channel chunk c0): a warp reads the TMEM
    // lane quarter warp % 4, so the two warps of a quarter split the COLUMNS -- group g owns channels [16 g, 16 g + 16)
    // of the chunk (the stage profile showed the single-group epilogue as the pace of the transposed stride-2 form, 83 %
    // busy, and the busiest role of the stride-2 gather form).  Thread m owns accumulator row (pixel) m.  The chunk goes
    // through a 128B-swizzled staging tile [128 pixels][32 channels]; one thread hands it to the TMA engine (a box
    // (32, 8, img, rows) of the (c, w, n, h) view of y: elements outside the tensor are clipped), double-buffered.
    const int grp = warp >= 16 ? 1 : 0;
    const int wq = warp & 3;
    int ab = 0;
    uint32_t pacc = 0;
    const int m = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    unsigned char* my_row0 = out_smem + (size_t)m * 128;
    const int sw = m & 7;
    const int epi = p.epi;
    const bool issuer = (tid == 0);
    // pixel of this thread inside a tile: group g = m / 8 = row * img + slot, column m % 8
    const int trow = (m >> 3) / p.img, tslot = (m >> 3) - trow * p.img, tcol = m & 7;
    const float inv_nt = 1.0f / (float)p.nt;
    uint32_t seq = 0;
    TC_PROF_DECL
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      const int img0 = (t / p.tiles_h) * p.img;
      TC_WAIT(&acc_full[ab], pacc, 0);
      tc::tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < G::NACC; ++a) {
        float rscale = 1.0f;
        float v[16];
        bool have_v = false;
        // this thread's 16 channels of chunk c0 (hi / lo column blocks of the cat form with one TMEM round trip)
        auto load_half = [&](int c0) {
          uint32_t pk[16], p2[16];
          const uint32_t col = tmem_base + lane_base + (uint32_t)(ab * acc_cols + a * acc_w + c0 + 16 * grp);
          tc::tmem_ld16_issue(col, pk);
          if (CAT) {
            tc::tmem_ld16_issue(col + (uint32_t)p.nt, p2);
            tc::tmem_ld_wait(pk, p2);
          } else {
            tc::tmem_ld_wait(pk);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(pk[j]) + (CAT ? __uint_as_float(p2[j]) : 0.0f);
        };
        if (epi == TC_EPI_PNF) {
          // mean square of the activated outputs of this pixel over ALL channels: each group sums its halves of every
          // chunk, the two exchange partial sums through shared memory
          float ss = 0.0f;
#pragma unroll 1
          for (int c0 = 0; c0 < p.nt; c0 += 32) {
            load_half(c0);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float o = gs_lrelu(fmaf(v[j], p.alpha, bias_s[c0 + 16 * grp + j]));
              ss = fmaf(o, o, ss);
            }
          }
          have_v = (p.nt == 32);
          ss_x[grp][m] = ss;
          asm volatile("bar.sync 3, 256;" ::: "memory");
          ss += ss_x[grp ^ 1][m];                 // rewritten only after the staging barrier of this accumulator's chunks
          rscale = 1.0f / sqrtf(ss * inv_nt + p.eps);
          const int pn = img0 + tslot;
          if (grp == 0 && pn < p.n_img) {
            int py = th_ * p.rows + trow, px = tw_ * 8 + tcol;
            if (FORM == TC_T2) { py = 2 * py + (a >> 1); px = 2 * px + (a & 1); }
            p.rvec[((size_t)pn * p.h_out + py) * p.w_out + px] = rscale;
          }
        }
#pragma unroll 1
        for (int c0 = 0; c0 < p.nt; c0 += 32, ++seq) {
          if (!have_v) load_half(c0);
          have_v = false;
          const uint32_t buf = seq & 1u;
          unsigned char* row = my_row0 + (size_t)buf * 16384;
          if (epi == TC_EPI_MASK) {
            const uint32_t aslot = seq % (uint32_t)p.aux_k, aph = (seq / (uint32_t)p.aux_k) & 1u;
            TC_WAIT(&aux_full[aslot], aph, 2);
            const unsigned char* arow = aux_smem + (size_t)aslot * 16384 + (size_t)m * 128;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int ch16 = 4 * grp + (j >> 2);
              const float4 ax = *reinterpret_cast<const float4*>(arow + ((ch16 ^ sw) << 4));
              float4 o;
              o.x = v[j + 0] * p.alpha * gs_lrelu_slope(ax.x); o.y = v[j + 1] * p.alpha * gs_lrelu_slope(ax.y);
              o.z = v[j + 2] * p.alpha * gs_lrelu_slope(ax.z); o.w = v[j + 3] * p.alpha * gs_lrelu_slope(ax.w);
              *reinterpret_cast<float4*>(row + ((ch16 ^ sw) << 4)) = o;
            }
            tc::mbar_arrive(&aux_empty[aslot]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int ch16 = 4 * grp + (j >> 2);
              const float4 bv = *reinterpret_cast<const float4*>(&bias_s[c0 + 16 * grp + j]);
              float4 o;
              o.x = fmaf(v[j + 0], p.alpha, bv.x); o.y = fmaf(v[j + 1], p.alpha, bv.y);
              o.z = fmaf(v[j + 2], p.alpha, bv.z); o.w = fmaf(v[j + 3], p.alpha, bv.w);
              if (p.act == 1 && p.ksplit == 1) { o.x = gs_lrelu(o.x); o.y = gs_lrelu(o.y); o.z = gs_lrelu(o.z); o.w = gs_lrelu(o.w); }
              if (epi == TC_EPI_PNF) { o.x *= rscale; o.y *= rscale; o.z *= rscale; o.w *= rscale; }
              *reinterpret_cast<float4*>(row + ((ch16 ^ sw) << 4)) = o;
            }
          }
          tc::fence_proxy_async();
          // staging tile seq & 1 was last read by the store of chunk seq - 2, which the issuer waited for before the
          // barrier of chunk seq - 1; here it waits for the store of chunk seq - 1 (its tile is rewritten by chunk seq + 1)
          TC_PROF_BEGIN(1)
          if (issuer) tc::bulk_wait_read<0>();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          TC_PROF_END(1)
          if (issuer) {
            if (p.ksplit > 1) tc::tma_reduce_add_4d(&tmy.m[a], out_smem + (size_t)buf * 16384, n0 + c0, tw_ * 8, img0, th_ * p.rows);
            else tc::tma_store_4d(&tmy.m[a], out_smem + (size_t)buf * 16384, n0 + c0, tw_ * 8, img0, th_ * p.rows);
            tc::bulk_commit();
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[ab]);
      if (++ab == p.nbuf) { ab = 0; pacc ^= 1u; }
    }
    TC_PROF_FLUSH(p.prof, 19, tid == 0);
    if (issuer) tc::bulk_wait<0>();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP0) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
