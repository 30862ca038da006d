// tcgen05 / TMEM implicit-GEMM 3x3 convolutions for sm_100a (NHWC fp32 in HBM, bf16x3 on the tensor cores).
//
// One CTA owns 128 output pixels (M = 128 accumulator rows = TMEM lanes: 16 groups of 8 consecutive
// columns) and NT <= 256 output channels (accumulator columns).  Data path per (tile, chunk of KC input
// channels):
//
//   HBM --TMA tiled load, fp32 halo box, hardware zero fill outside the image--> raw ring (128B / 64B swizzle)
//       --8 converter warps: fp32 -> (bf16 hi, bf16 lo)--> operand ring: 16-byte channel vectors, pixel-major
//         ("SWIZZLE_NONE core matrices": plane q = channels 8q..8q+7, 16 bytes per pixel)
//       --tcgen05.mma, one elected thread--> TMEM accumulators --4 epilogue warps--> alpha, bias, leaky-relu --> HBM
//
// In the operand layout every filter tap is just a different START ADDRESS of the same staged tile (group
// stride fixed), so the nine taps cost nine descriptor pairs, not nine loads.  Each 16-channel K slice
// issues three MMAs (hi*hi + hi*lo + lo*hi); accumulation is fp32 in TMEM (the accumulator truncates:
// ~1e-5 relative at K = 2304, tools/tc_precision.py).  Weights are pre-split into the same layout by
// conv_tc_prep_kernel, contiguous per (output-channel tile, chunk), and streamed with 1-D bulk copies in
// groups of `tps` taps (or loaded once and kept when they fit).
// Warp roles: 0-3 epilogue, 4-11 converters, 12 TMA halo loads, 13 weight copies, 14 MMA issue + TMEM
// allocation.  Rings: raw (TMA -> converters), A (converters -> MMA), B (weights), accumulators (per tile).
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
  const __nv_bfloat16* wprep;     // [ndim/nt][kdim/KC][9][2][KC/8][nt][8]
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
  int tps;                        // taps per weight stage (9, 3 or 1)
  int b_resident;                 // the weights of this CTA's channel tile are loaded once and kept
  int nbuf;                       // accumulator buffers (1 or 2)
  int tmem_cols;                  // power of two >= nbuf * NACC * nt
  uint32_t raw_slot_bytes;        // 1024-byte aligned
};

constexpr int TC_THREADS = 480;
constexpr int TC_CONV_WARPS = 8;
constexpr int TC_MAX_STAGES = 4;       // raw and operand rings
constexpr int TC_MAX_BSTAGES = 24;     // weight ring (all (chunk, tap group) blocks when resident)

// Weight pre-pass: W_eff[tap][k][n] (k = contraction channel, n = output channel) -> bf16 split blocks
// [n / nt][k / KC][tap][split][q][n % nt][e], value index k = KC*kc + 8*q + e.  w_is_kn: weight memory is
// [tap][k][n] (else [tap][n][k]); flip: use tap 8 - t (180-degree rotation).
template <int KC>
__global__ void conv_tc_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim, int ndim,
                                    int nt, int w_is_kn, int flip) {
  constexpr int Q = KC / 8;
  const size_t total = (size_t)9 * kdim * ndim;
  const int nchunks = kdim / KC;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
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
    out[(blk + q) * plane + inner] = hi;
    out[(blk + Q + q) * plane + inner] = lo;
  }
}

template <int FORM, int KC>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmx, const TcParams p) {
  using G = TcGeo<FORM>;
  constexpr int Q = KC / 8;               // 16-byte channel planes per split term
  extern __shared__ unsigned char tc_smem_raw[];
  __shared__ uint64_t raw_full[TC_MAX_STAGES], raw_empty[TC_MAX_STAGES], a_full[TC_MAX_STAGES], a_empty[TC_MAX_STAGES];
  __shared__ uint64_t b_full[TC_MAX_BSTAGES], b_empty[TC_MAX_BSTAGES];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t plane_a = (uint32_t)p.pix * 16u;
  const uint32_t plane_b = (uint32_t)p.nt * 16u;
  const uint32_t a_stage_bytes = (uint32_t)(2 * Q) * plane_a;   // [split][q] planes
  const uint32_t b_tap_bytes = (uint32_t)(2 * Q) * plane_b;
  const uint32_t b_stage_bytes = (uint32_t)p.tps * b_tap_bytes;
  // the swizzled TMA destination needs 1024-byte alignment
  unsigned char* tc_smem = tc_smem_raw + ((1024u - (tc::smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char* raw_smem = tc_smem;
  unsigned char* a_smem = raw_smem + (size_t)p.ds * p.raw_slot_bytes;
  unsigned char* b_smem = a_smem + (size_t)p.sa * a_stage_bytes;
  uint16_t* dst_tab = reinterpret_cast<uint16_t*>(b_smem + (size_t)p.sb * b_stage_bytes);   // TC_C2 only
  const int nchunks = p.kdim / KC;
  const int n0 = blockIdx.y * p.nt;       // first output channel of this CTA

  if (tid == 0) {
    for (int s = 0; s < p.ds; ++s) { tc::mbar_init(&raw_full[s], 1); tc::mbar_init(&raw_empty[s], TC_CONV_WARPS * 32); }
    for (int s = 0; s < p.sa; ++s) { tc::mbar_init(&a_full[s], TC_CONV_WARPS * 32); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.sb; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 128); }
    tc::mbar_fence_init();
  }
  if (FORM == TC_C2) {
    // raw pixel (row hr, image slot, column hc) -> staged position: rows in pairs, columns split by parity
    for (int ps = tid; ps < p.rpix; ps += TC_THREADS) {
      const int hc = ps % 17, u = ps / 17;
      const int slot = u % p.img, hr = u / p.img;
      dst_tab[ps] = (uint16_t)(((hr >> 1) * p.img + slot) * 34 + (hr & 1) * 17 + (hc & 1) * 9 + (hc >> 1));
    }
  }
  if (warp == 14) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp == 12 && lane == 0) tc::prefetch_tmap(&tmx);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int acc_cols = G::NACC * p.nt;   // columns per accumulator buffer

  if (warp >= 4 && warp < 12) {
    // ============================== fp32 -> bf16 hi/lo split ===========================================
    // item = (raw pixel, 8-channel group q): 32 bytes in, 16 + 16 bytes out.  A quarter warp reads 8
    // consecutive pixels of one q: the TMA swizzle makes that conflict-free, the stores are contiguous.
    const int cw = warp - 4;
    const int q = cw % Q;
    constexpr int PSTEP = 32 * (TC_CONV_WARPS / Q);
    const int p_first = (cw / Q) * 32 + lane;
    int stage = 0, rs = 0;
    uint32_t aph = 0, rph = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int kc = 0; kc < nchunks; ++kc) {
        tc::mbar_wait(&raw_full[rs], rph);
        tc::mbar_wait(&a_empty[stage], aph ^ 1u);
        const unsigned char* raw = raw_smem + (size_t)rs * p.raw_slot_bytes;
        unsigned char* st = a_smem + (size_t)stage * a_stage_bytes;
        unsigned char* st_hi = st + (size_t)q * plane_a;
        unsigned char* st_lo = st + (size_t)(Q + q) * plane_a;
#pragma unroll 2
        for (int ps = p_first; ps < p.rpix; ps += PSTEP) {
          const int sw = (KC == 32) ? (ps & 7) : ((ps >> 1) & 3);
          const unsigned char* row = raw + (size_t)ps * (KC * 4);
          const float4 v0 = *reinterpret_cast<const float4*>(row + (((2 * q) ^ sw) << 4));
          const float4 v1 = *reinterpret_cast<const float4*>(row + (((2 * q + 1) ^ sw) << 4));
          uint4 h4, l4;
          tc::split2_bf16(v0.x, v0.y, h4.x, l4.x);
          tc::split2_bf16(v0.z, v0.w, h4.y, l4.y);
          tc::split2_bf16(v1.x, v1.y, h4.z, l4.z);
          tc::split2_bf16(v1.z, v1.w, h4.w, l4.w);
          const int d = (FORM == TC_C2) ? (int)dst_tab[ps] : ps;
          *reinterpret_cast<uint4*>(st_hi + (size_t)d * 16) = h4;
          *reinterpret_cast<uint4*>(st_lo + (size_t)d * 16) = l4;
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&a_full[stage]);
        tc::mbar_arrive(&raw_empty[rs]);
        if (++stage == p.sa) { stage = 0; aph ^= 1u; }
        if (++rs == p.ds) { rs = 0; rph ^= 1u; }
      }
    }
  } else if (warp == 12) {
    // ============================== halo tiles: one TMA box per (tile, chunk) ===========================
    if (lane == 0) {
      int rs = 0;
      uint32_t rph = 0;
      const uint32_t box_bytes = (uint32_t)p.rpix * (KC * 4u);
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int t = tile;
        const int tw_ = t % p.tiles_w;
        t /= p.tiles_w;
        const int th_ = t % p.tiles_h;
        const int img0 = (t / p.tiles_h) * p.img;
        const int w0 = (FORM == TC_C2) ? tw_ * 16 : tw_ * 8 - 1;
        const int h0 = (FORM == TC_C2) ? th_ * 32 : th_ * 16 - 1;
        for (int kc = 0; kc < nchunks; ++kc) {
          tc::mbar_wait(&raw_empty[rs], rph ^ 1u);
          tc::mbar_arrive_expect_tx(&raw_full[rs], box_bytes);
          tc::tma_load_4d(raw_smem + (size_t)rs * p.raw_slot_bytes, &tmx, kc * KC, w0, img0, h0, &raw_full[rs]);
          if (++rs == p.ds) { rs = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp == 13) {
    // ============================== weight blocks: one bulk copy per (chunk, tap group) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int groups = 9 / p.tps;
      const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wprep) + (size_t)blockIdx.y * nchunks * 9 * b_tap_bytes;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        if (p.b_resident && tile != (int)blockIdx.x) break;     // resident weights: one pass fills every stage
        for (int kc = 0; kc < nchunks; ++kc) {
          for (int gi = 0; gi < groups; ++gi) {
            if (!p.b_resident) tc::mbar_wait(&b_empty[stage], phase ^ 1u);
            tc::mbar_arrive_expect_tx(&b_full[stage], b_stage_bytes);
            tc::bulk_g2s(b_smem + (size_t)stage * b_stage_bytes, wsrc + (size_t)(kc * groups + gi) * b_stage_bytes, b_stage_bytes,
                         &b_full[stage]);
            if (++stage == p.sb) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 14) {
    // ============================== MMA issue ==========================================================
    // One thread feeds the tensor core.  With N = 32 an MMA is only ~16 cycles of tensor work, so the
    // issue loop is kept to an add or two per MMA: descriptors are a per-stage base (start address in
    // 16-byte units in the low word) plus tap / K-slice / split offsets.  The WHOLE warp runs this control
    // flow (so the compiler keeps descriptors in uniform registers); one elected lane issues.
    const uint32_t idesc = tc::idesc_bf16_f32(p.nt, 0, 0);
    const uint64_t a_desc0 = tc::smem_desc(tc::smem_u32(a_smem), plane_a, G::GSTRIDE * 16u);
    const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(b_smem), plane_b, 128u);
    const uint32_t a_stage16 = a_stage_bytes >> 4, b_stage16 = b_stage_bytes >> 4, b_tap16 = b_tap_bytes >> 4;
    const uint32_t plane_a16 = plane_a >> 4, plane_b16 = plane_b >> 4;
    int sa = 0, sb = 0, ab = 0, tin = 0;             // tin: tap index inside the current weight stage
    uint32_t pa = 0, pb = 0, pacc = 0;
    bool b_ready = false;                            // resident weights: waited for once
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      tc::mbar_wait(&acc_empty[ab], pacc ^ 1u);
      tc::tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(ab * acc_cols);
      for (int kc = 0; kc < nchunks; ++kc) {
        tc::mbar_wait(&a_full[sa], pa);
        tc::tc_fence_after();
        const uint64_t a_base = a_desc0 + (uint64_t)((uint32_t)sa * a_stage16);
        const uint32_t acc_rest = (kc > 0) ? 1u : 0u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if (tin == 0 && !b_ready) {
            tc::mbar_wait(&b_full[sb], pb);
            tc::tc_fence_after();
          }
          const uint64_t b_base = b_desc0 + (uint64_t)((uint32_t)sb * b_stage16 + (uint32_t)tin * b_tap16);
          const int acc = G::tap_acc(tap);
          const uint32_t d = d0 + (uint32_t)(acc * p.nt);
          const uint64_t a_tap = a_base + (uint64_t)(uint32_t)p.tap_off[tap];
          // the first tap that touches an accumulator overwrites it on the first channel chunk
          const bool opens = (G::NACC == 1) ? (tap == 0) : (tap == 0 || tap == 1 || tap == 3 || tap == 4);
          const bool last_in_stage = (tin == p.tps - 1);
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
              constexpr int PA[3] = {0, 0, 1};
              constexpr int PB[3] = {0, 1, 0};
#pragma unroll
              for (int pr = 0; pr < 3; ++pr) {
                const uint64_t da = a_tap + (uint64_t)((uint32_t)(PA[pr] * Q + 2 * ks) * plane_a16);
                const uint64_t db = b_base + (uint64_t)((uint32_t)(PB[pr] * Q + 2 * ks) * plane_b16);
                const uint32_t accum = (opens && ks == 0 && pr == 0) ? acc_rest : 1u;
                tc::mma_bf16(d, da, db, idesc, accum);
              }
            }
            if (last_in_stage && !p.b_resident) tc::mma_commit(&b_empty[sb]);
            if (tap == 8) {
              tc::mma_commit(&a_empty[sa]);
              if (kc == nchunks - 1) tc::mma_commit(&acc_full[ab]);
            }
          }
          __syncwarp();
          if (last_in_stage) {
            tin = 0;
            if (++sb == p.sb) { sb = 0; pb ^= 1u; }
          } else {
            ++tin;
          }
        }
        if (++sa == p.sa) { sa = 0; pa ^= 1u; }
      }
      if (++ab == p.nbuf) { ab = 0; pacc ^= 1u; }
      if (p.b_resident) b_ready = true;
    }
  } else if (warp < 4) {
    // ============================== epilogue: TMEM -> alpha, bias, leaky-relu -> HBM =====================
    int ab = 0;
    uint32_t pacc = 0;
    const int m = warp * 32 + lane;
    const int g = m >> 3, i = m & 7;
    const int row = g / p.img, slot = g % p.img;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      const int n = (t / p.tiles_h) * p.img + slot;
      tc::mbar_wait(&acc_full[ab], pacc);
      tc::tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < G::NACC; ++a) {
        int oy, ox;
        if (FORM == TC_T2) { oy = 2 * (th_ * 16 + row) + (a >> 1); ox = 2 * (tw_ * 8 + i) + (a & 1); }
        else { oy = th_ * 16 + row; ox = tw_ * 8 + i; }
        const bool in_range = (n < p.n_img) && (oy < p.h_out) && (ox < p.w_out);
        float* dst = p.y + (((size_t)n * p.h_out + oy) * p.w_out + ox) * p.ndim + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < p.nt; c0 += 32) {
          float v[32];
          tc::tmem_ld32(tmem_base + lane_base + (uint32_t)(ab * acc_cols + a * p.nt + c0), v);
          if (in_range) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o;
              float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
              o.x = fmaf(v[j + 0], p.alpha, bv.x); o.y = fmaf(v[j + 1], p.alpha, bv.y);
              o.z = fmaf(v[j + 2], p.alpha, bv.z); o.w = fmaf(v[j + 3], p.alpha, bv.w);
              if (p.act == 1) { o.x = gs_lrelu(o.x); o.y = gs_lrelu(o.y); o.z = gs_lrelu(o.z); o.w = gs_lrelu(o.w); }
              *reinterpret_cast<float4*>(dst + c0 + j) = o;
            }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[ab]);
      if (++ab == p.nbuf) { ab = 0; pacc ^= 1u; }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 14) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
