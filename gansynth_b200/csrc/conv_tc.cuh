// tcgen05 / TMEM implicit-GEMM 3x3 convolutions for sm_100a (NHWC fp32 in HBM, bf16x3 on the tensor cores).
//
// One CTA owns a 16x8 pixel tile (M = 128 accumulator rows = TMEM lanes) and ALL output channels
// (N = Cout <= 256 accumulator columns).  For each chunk of 32 input channels the halo of the tile is read
// from HBM once, split fp32 -> (bf16 hi, bf16 lo) and stored in shared memory as 16-byte channel vectors,
// pixel-major ("SWIZZLE_NONE core matrices": plane q = channels 8q..8q+7, 16 bytes per pixel).  In that
// layout every filter tap is just a different START ADDRESS / group stride of the same staged tile, so the
// nine taps cost nine descriptor pairs, not nine loads.  Each 16-channel K slice issues three MMAs
// (hi*hi + hi*lo + lo*hi; NSPLIT = 2: products carry ~17 mantissa bits) or, for layer forwards whose sign
// decides the leaky-relu mask, six MMAs of a 3-way split (h,m,l; NSPLIT = 3: ~25 bits, fp32-grade);
// accumulation is fp32 in TMEM.
// Weights are pre-split into the same layout by conv_tc_prep_kernel and streamed per tap with 1-D bulk
// copies.  Warp roles: 0-3 epilogue (TMEM -> bias / leaky-relu -> HBM), 4-7 halo load + split, 8 weight
// copies, 9 MMA issue + TMEM allocation.  Rings: A (per channel chunk), B (per tap), accumulators (per tile).
//
// Forms (template FORM):
//   TC_C1  gather conv stride 1 (forward of conv2d; dgrad of stride 1 with flipped/transposed weights)
//   TC_C2  gather conv stride 2 (forward of the down-scaling conv; dgrad of conv2d_transpose)
//   TC_T2  transposed conv stride 2 (forward of conv2d_transpose; dgrad of the down-scaling conv):
//          four sub-pixel phases = four accumulators fed by 4/2/2/1 taps of one 17x9 halo
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

enum { TC_C1 = 0, TC_C2 = 1, TC_T2 = 2 };

template <int FORM>
struct TcGeo;
template <>
struct TcGeo<TC_C1> {
  static constexpr int P = 18 * 10, GSTRIDE = 10, NACC = 1;
  __device__ static int tap_off(int t) { return (t / 3) * 10 + (t % 3); }
  __device__ static int tap_acc(int) { return 0; }
};
template <>
struct TcGeo<TC_C2> {
  // 33 input rows x 17 input columns, columns de-interleaved by parity: [row][9 even | 8 odd]
  static constexpr int P = 33 * 17, GSTRIDE = 34, NACC = 1;
  __device__ static int tap_off(int t) {
    int kh = t / 3, kw = t % 3;
    return kh * 17 + (kw & 1) * 9 + (kw >> 1);
  }
  __device__ static int tap_acc(int) { return 0; }
};
template <>
struct TcGeo<TC_T2> {
  // 17 rows (i0-1 .. i0+15) x 9 columns (j0-1 .. j0+7) of the small side
  static constexpr int P = 17 * 9, GSTRIDE = 9, NACC = 4;
  __device__ static int tap_off(int t) {
    int kh = t / 3, kw = t % 3;
    return ((kh == 2) ? 0 : 1) * 9 + ((kw == 2) ? 0 : 1);
  }
  __device__ static int tap_acc(int t) { return ((t / 3) == 1 ? 2 : 0) + ((t % 3) == 1 ? 1 : 0); }
};

struct TcParams {
  const float* x;                 // A-side activations [n, h_in, w_in, kdim]
  const __nv_bfloat16* wprep;     // [kdim/KC][9][NSPLIT][KC/8][ndim][8]
  const float* bias;              // [ndim] or null
  float* y;                       // [n, h_out, w_out, ndim]
  int n_img, h_in, w_in, h_out, w_out, kdim, ndim;
  float alpha;
  int act;
  int tiles_h, tiles_w, ntiles;
  int sa, sb;                     // ring depths
  int nbuf;                       // accumulator buffers (1 or 2)
  int tmem_cols;                  // power of two >= nbuf * NACC * ndim
};

constexpr int TC_THREADS = 320;
constexpr int TC_MAX_STAGES = 4;

// fp32 -> NSPLIT bf16 terms whose sum reproduces x to 2^-17 (2 terms) / 2^-25 (3 terms) relative
template <int NSPLIT>
__device__ __forceinline__ void tc_split(float x, __nv_bfloat16 (&t)[NSPLIT]) {
  float r = x;
#pragma unroll
  for (int s = 0; s < NSPLIT; ++s) {
    t[s] = __float2bfloat16_rn(r);
    r -= __bfloat162float(t[s]);
  }
}

// Weight pre-pass: W_eff[tap][k][n] (k = contraction channel, n = output channel) -> bf16 split blocks
// [k/KC][tap][split][q][n][e], value index k = KC*kc + 8*q + e.  w_is_kn / flip as in gs_load_b_tile.
template <int NSPLIT, int KC>
__global__ void conv_tc_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim, int ndim,
                                    int w_is_kn, int flip) {
  constexpr int Q = KC / 8;
  const size_t total = (size_t)9 * kdim * ndim;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int e = (int)(i % 8);
    size_t r = i / 8;
    int n = (int)(r % ndim);
    r /= ndim;
    int q = (int)(r % Q);
    r /= Q;
    int tap = (int)(r % 9);
    int kc = (int)(r / 9);
    int k = kc * KC + q * 8 + e;
    int st = flip ? 8 - tap : tap;
    float v = w_is_kn ? w[((size_t)st * kdim + k) * ndim + n] : w[((size_t)st * ndim + n) * kdim + k];
    __nv_bfloat16 t[NSPLIT];
    tc_split<NSPLIT>(v, t);
    size_t blk = ((size_t)kc * 9 + tap) * NSPLIT;
    size_t plane = (size_t)Q * ndim * 8;
    size_t inner = ((size_t)q * ndim + n) * 8 + e;
#pragma unroll
    for (int sp = 0; sp < NSPLIT; ++sp) out[(blk + sp) * plane + inner] = t[sp];
  }
}

template <int FORM, int NSPLIT, int KC>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcParams p) {
  using G = TcGeo<FORM>;
  constexpr int Q = KC / 8;               // 16-byte channel planes per split term
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ uint64_t a_full[TC_MAX_STAGES], a_empty[TC_MAX_STAGES], b_full[TC_MAX_STAGES], b_empty[TC_MAX_STAGES];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t plane_a = G::P * 16u;
  const uint32_t plane_b = (uint32_t)p.ndim * 16u;
  const uint32_t a_stage_bytes = (uint32_t)(NSPLIT * Q) * plane_a;   // [split][q] planes
  const uint32_t b_stage_bytes = (uint32_t)(NSPLIT * Q) * plane_b;
  unsigned char* a_smem = tc_smem;
  unsigned char* b_smem = tc_smem + (size_t)p.sa * a_stage_bytes;
  const int nchunks = p.kdim / KC;

  if (tid == 0) {
    for (int s = 0; s < p.sa; ++s) { tc::mbar_init(&a_full[s], 128); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.sb; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 128); }
    tc::mbar_fence_init();
  }
  if (warp == 9) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int acc_cols = G::NACC * p.ndim;   // columns per accumulator buffer

  if (warp >= 4 && warp < 8) {
    // ============================== halo load + fp32 -> bf16 hi/lo split ==============================
    const int ct = tid - 128;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      const int n = t / p.tiles_h;
      int r0, c0;   // input coordinates of halo element (0, 0)
      if (FORM == TC_C1) { r0 = th_ * 16 - 1; c0 = tw_ * 8 - 1; }
      else if (FORM == TC_C2) { r0 = th_ * 32; c0 = tw_ * 16; }
      else { r0 = th_ * 16 - 1; c0 = tw_ * 8 - 1; }
      for (int kc = 0; kc < nchunks; ++kc) {
        tc::mbar_wait(&a_empty[stage], phase ^ 1u);
        unsigned char* st = a_smem + (size_t)stage * a_stage_bytes;
        for (int ps = ct; ps < G::P; ps += 128) {
          int hr, hc;
          if (FORM == TC_C1) { hr = ps / 10; hc = ps % 10; }
          else if (FORM == TC_C2) { hr = ps / 17; int rem = ps % 17; int par = rem >= 9; hc = 2 * (rem - 9 * par) + par; }
          else { hr = ps / 9; hc = ps % 9; }
          const int ih = r0 + hr, iw = c0 + hc;
          float4 v[2 * Q];
          if (ih >= 0 && ih < p.h_in && iw >= 0 && iw < p.w_in) {
            const float4* src = reinterpret_cast<const float4*>(p.x + (((size_t)n * p.h_in + ih) * p.w_in + iw) * p.kdim + kc * KC);
#pragma unroll
            for (int j = 0; j < 2 * Q; ++j) v[j] = __ldg(src + j);
          } else {
#pragma unroll
            for (int j = 0; j < 2 * Q; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float f[8] = {v[2 * q].x, v[2 * q].y, v[2 * q].z, v[2 * q].w, v[2 * q + 1].x, v[2 * q + 1].y, v[2 * q + 1].z, v[2 * q + 1].w};
            __nv_bfloat16 t[8][NSPLIT];
#pragma unroll
            for (int e = 0; e < 8; ++e) tc_split<NSPLIT>(f[e], t[e]);
#pragma unroll
            for (int sp = 0; sp < NSPLIT; ++sp) {
              uint4 o;
              o.x = tc::pack_bf16(t[0][sp], t[1][sp]); o.y = tc::pack_bf16(t[2][sp], t[3][sp]);
              o.z = tc::pack_bf16(t[4][sp], t[5][sp]); o.w = tc::pack_bf16(t[6][sp], t[7][sp]);
              *reinterpret_cast<uint4*>(st + (size_t)(sp * Q + q) * plane_a + (size_t)ps * 16) = o;
            }
          }
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&a_full[stage]);
        if (++stage == p.sa) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 8) {
    // ============================== weight blocks: one bulk copy per (chunk, tap) ======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nchunks; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            tc::mbar_wait(&b_empty[stage], phase ^ 1u);
            tc::mbar_arrive_expect_tx(&b_full[stage], b_stage_bytes);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wprep) + ((size_t)kc * 9 + tap) * b_stage_bytes;
            tc::bulk_g2s(b_smem + (size_t)stage * b_stage_bytes, src, b_stage_bytes, &b_full[stage]);
            if (++stage == p.sb) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ============================== MMA issue ==========================================================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_bf16_f32(p.ndim, 0, 0);
      int sa = 0, sb = 0, ab = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        tc::mbar_wait(&acc_empty[ab], pacc ^ 1u);
        tc::tc_fence_after();
        uint32_t started = 0;   // bit a: accumulator a already holds a partial sum for this tile
        for (int kc = 0; kc < nchunks; ++kc) {
          tc::mbar_wait(&a_full[sa], pa);
          tc::tc_fence_after();
          const uint32_t a0 = tc::smem_u32(a_smem + (size_t)sa * a_stage_bytes);
          for (int tap = 0; tap < 9; ++tap) {
            tc::mbar_wait(&b_full[sb], pb);
            tc::tc_fence_after();
            const uint32_t b0 = tc::smem_u32(b_smem + (size_t)sb * b_stage_bytes);
            const int acc = G::tap_acc(tap);
            const uint32_t d = tmem_base + (uint32_t)(ab * acc_cols + acc * p.ndim);
            const uint32_t aoff = (uint32_t)G::tap_off(tap) * 16u;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
              // split products kept: (0,0) (0,1) (1,0) [+ (0,2) (2,0) (1,1) for the 3-way split]
              constexpr int NPROD = (NSPLIT == 2) ? 3 : 6;
              constexpr int PA[6] = {0, 0, 1, 0, 2, 1};
              constexpr int PB[6] = {0, 1, 0, 2, 0, 1};
#pragma unroll
              for (int pr = 0; pr < NPROD; ++pr) {
                const uint64_t da = tc::smem_desc(a0 + (uint32_t)(PA[pr] * Q + 2 * ks) * plane_a + aoff, plane_a, G::GSTRIDE * 16u);
                const uint64_t db = tc::smem_desc(b0 + (uint32_t)(PB[pr] * Q + 2 * ks) * plane_b, plane_b, 128u);
                tc::mma_bf16(d, da, db, idesc, (pr == 0) ? ((started >> acc) & 1u) : 1u);
              }
              started |= 1u << acc;
            }
            tc::mma_commit(&b_empty[sb]);
            if (++sb == p.sb) { sb = 0; pb ^= 1u; }
          }
          tc::mma_commit(&a_empty[sa]);
          if (++sa == p.sa) { sa = 0; pa ^= 1u; }
        }
        tc::mma_commit(&acc_full[ab]);
        if (++ab == p.nbuf) { ab = 0; pacc ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue: TMEM -> alpha, bias, leaky-relu -> HBM =====================
    int ab = 0;
    uint32_t pacc = 0;
    const int m = warp * 32 + lane;
    const int g = m >> 3, i = m & 7;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      const int n = t / p.tiles_h;
      tc::mbar_wait(&acc_full[ab], pacc);
      tc::tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < G::NACC; ++a) {
        int oy, ox;
        if (FORM == TC_T2) { oy = 2 * (th_ * 16 + g) + (a >> 1); ox = 2 * (tw_ * 8 + i) + (a & 1); }
        else { oy = th_ * 16 + g; ox = tw_ * 8 + i; }
        float* dst = p.y + (((size_t)n * p.h_out + oy) * p.w_out + ox) * p.ndim;
        const bool in_range = (oy < p.h_out) && (ox < p.w_out);
#pragma unroll 1
        for (int c0 = 0; c0 < p.ndim; c0 += 32) {
          float v[32];
          tc::tmem_ld32(tmem_base + lane_base + (uint32_t)(ab * acc_cols + a * p.ndim + c0), v);
          const int nvalid = p.ndim - c0;   // ndim is a multiple of 16: a chunk holds 32 or 16 valid columns
          if (in_range) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < nvalid) {
                float4 o;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                o.x = fmaf(v[j + 0], p.alpha, bv.x); o.y = fmaf(v[j + 1], p.alpha, bv.y);
                o.z = fmaf(v[j + 2], p.alpha, bv.z); o.w = fmaf(v[j + 3], p.alpha, bv.w);
                if (p.act == 1) { o.x = gs_lrelu(o.x); o.y = gs_lrelu(o.y); o.z = gs_lrelu(o.z); o.w = gs_lrelu(o.w); }
                *reinterpret_cast<float4*>(dst + c0 + j) = o;
              }
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
  if (warp == 9) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
