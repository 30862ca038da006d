// tcgen05 / TMEM implicit-GEMM 3x3 convolutions for sm_100a (NHWC fp32 in HBM, bf16x3 on the tensor cores).
//
// One CTA owns 128 output pixels (M = 128 accumulator rows = TMEM lanes: 16 groups of 8 consecutive
// columns) and NT <= 256 output channels (accumulator columns).  For each chunk of 32 input channels the
// halo of the tile is read from HBM once, split fp32 -> (bf16 hi, bf16 lo) and stored in shared memory as
// 16-byte channel vectors, pixel-major ("SWIZZLE_NONE core matrices": plane q = channels 8q..8q+7,
// 16 bytes per pixel).  In that layout every filter tap is just a different START ADDRESS of the same
// staged tile (group stride fixed), so the nine taps cost nine descriptor pairs, not nine loads.  Each
// 16-channel K slice issues three MMAs (hi*hi + hi*lo + lo*hi); accumulation is fp32 in TMEM (the
// accumulator truncates: ~1e-5 relative at K = 2304, tools/tc_precision.py).
// Weights are pre-split into the same layout by conv_tc_prep_kernel and streamed per tap with 1-D bulk
// copies (TMA engine).  Warp roles: 0-3 epilogue (TMEM -> alpha, bias, leaky-relu -> HBM), 4-7 halo load +
// split, 8 weight copies, 9 MMA issue + TMEM allocation.  Rings: A (per channel chunk), B (per tap),
// accumulators (per tile).
//
// Tiles.  Images with >= 16 rows: 16 rows x 8 columns of one image.  Smaller images (8, 4, 2 rows: the
// 256-channel low-resolution blocks) interleave IMG = 16 / rows images row by row -- group g = row * IMG
// + slot -- so the halo rows of different images never alias and the group stride stays uniform.
//
// Forms (template FORM):
//   TC_C1  gather conv stride 1 (forward of conv2d; dgrad of stride 1 with flipped/transposed weights)
//   TC_C2  gather conv stride 2 (forward of the down-scaling conv; dgrad of conv2d_transpose)
//   TC_T2  transposed conv stride 2 (forward of conv2d_transpose; dgrad of the down-scaling conv):
//          four sub-pixel phases = four accumulators fed by 4/2/2/1 taps of one halo
#pragma once
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
  // start offset (pixels) of tap t
  __host__ __device__ static constexpr int tap_off(int t, int img) {
    const int kh = t / 3, kw = t % 3;
    if (FORM == TC_C1) return kh * img * 10 + kw;
    if (FORM == TC_C2) return (kh == 0 ? 0 : kh == 1 ? 17 : img * 34) + (kw & 1) * 9 + (kw >> 1);
    return ((kh == 2) ? 0 : 1) * img * 9 + ((kw == 2) ? 0 : 1);
  }
};

struct TcParams {
  const float* x;                 // A-side activations [n, h_in, w_in, kdim]
  const __nv_bfloat16* wprep;     // [kdim/KC][9][NSPLIT][KC/8][ndim][8]
  const float* bias;              // [ndim] or null
  float* y;                       // [n, h_out, w_out, ndim]
  int n_img, h_in, w_in, h_out, w_out, kdim, ndim;
  float alpha;
  int act;
  int rows, img;                  // tile: `rows` rows of each of `img` interleaved images, 8 columns
  int pix;                        // staged pixels per tile
  int tiles_h, tiles_w, ntiles;   // ntiles = image groups * tiles_h * tiles_w
  int nt, n_tiles;                // output-channel tile (grid.y)
  int tap_off[9];
  int sa, sb;                     // ring depths
  int ds;                         // fp32 staging slots of the asynchronous halo prefetch (0: direct loads)
  int b_resident;                 // weights of this CTA's channel tile are loaded once and kept (sb = 9 * chunks)
  int nbuf;                       // accumulator buffers (1 or 2)
  int tmem_cols;                  // power of two >= nbuf * NACC * nt
};

constexpr int TC_THREADS = 320;
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_MAX_BSTAGES = 18;     // weight ring; 9 * chunks stages when the weights stay resident

// fp32 -> NSPLIT bf16 terms whose sum reproduces x to 2^-17 (2 terms) relative
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
  __shared__ uint64_t a_full[TC_MAX_STAGES], a_empty[TC_MAX_STAGES], b_full[TC_MAX_BSTAGES], b_empty[TC_MAX_BSTAGES];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t plane_a = (uint32_t)p.pix * 16u;
  const uint32_t plane_b = (uint32_t)p.nt * 16u;
  const uint32_t a_stage_bytes = (uint32_t)(NSPLIT * Q) * plane_a;   // [split][q] planes
  const uint32_t b_stage_bytes = (uint32_t)(NSPLIT * Q) * plane_b;
  unsigned char* a_smem = tc_smem;
  unsigned char* b_smem = tc_smem + (size_t)p.sa * a_stage_bytes;
  unsigned char* stg_smem = b_smem + (size_t)p.sb * b_stage_bytes;     // [ds][pix][KC] fp32, chunks XOR-swizzled
  const uint32_t stg_bytes = (uint32_t)p.pix * (KC * 4u);
  const int nchunks = p.kdim / KC;
  const int n0 = blockIdx.y * p.nt;       // first output channel of this CTA

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
  const int acc_cols = G::NACC * p.nt;   // columns per accumulator buffer

  if (warp >= 4 && warp < 8) {
    // ============================== halo load + fp32 -> bf16 hi/lo split ==============================
    // The loads are the latency-critical part (one HBM round trip per tile otherwise): they are issued
    // p.ds work items ahead as 16-byte cp.async copies into an fp32 staging ring (zero-filled outside the
    // image).  Every thread converts exactly the pixels it copied itself, so cp.async.wait_group is the
    // only synchronisation the staging ring needs.
    const int ct = tid - 128;
    constexpr int CH = KC / 4;                      // 16-byte chunks per pixel
    // bank-conflict-free chunk permutation of the staging rows (128-byte rows: 8 pixels x 8 chunks;
    // 64-byte rows: pixel pairs share a 128-byte line)
    auto swz = [](int ps) { return (CH == 8) ? (ps & 7) : ((ps >> 1) & 3); };
    auto coords = [&](int tile, int ps, int& n, int& ih, int& iw) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      const int img0 = (t / p.tiles_h) * p.img;
      int hr, hc, slot;
      if (FORM == TC_C1) { hc = ps % 10; int u = ps / 10; slot = u % p.img; hr = u / p.img; ih = th_ * 16 - 1 + hr; iw = tw_ * 8 - 1 + hc; }
      else if (FORM == TC_C2) {
        int rem = ps % 34, u = ps / 34;
        slot = u % p.img;
        hr = 2 * (u / p.img) + (rem >= 17);
        rem -= 17 * (rem >= 17);
        int par = rem >= 9;
        hc = 2 * (rem - 9 * par) + par;
        ih = th_ * 32 + hr; iw = tw_ * 16 + hc;
      } else { hc = ps % 9; int u = ps / 9; slot = u % p.img; hr = u / p.img; ih = th_ * 16 - 1 + hr; iw = tw_ * 8 - 1 + hc; }
      n = img0 + slot;
    };
    auto convert_store = [&](unsigned char* st, int ps, const float4 (&v)[CH]) {
      static_assert(NSPLIT == 2, "the packed split handles the two-term form");
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        uint4 h4, l4;
        tc::split2_bf16(v[2 * q].x, v[2 * q].y, h4.x, l4.x);
        tc::split2_bf16(v[2 * q].z, v[2 * q].w, h4.y, l4.y);
        tc::split2_bf16(v[2 * q + 1].x, v[2 * q + 1].y, h4.z, l4.z);
        tc::split2_bf16(v[2 * q + 1].z, v[2 * q + 1].w, h4.w, l4.w);
        *reinterpret_cast<uint4*>(st + (size_t)q * plane_a + (size_t)ps * 16) = h4;
        *reinterpret_cast<uint4*>(st + (size_t)(Q + q) * plane_a + (size_t)ps * 16) = l4;
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    if (p.ds == 0) {
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nchunks; ++kc) {
          tc::mbar_wait(&a_empty[stage], phase ^ 1u);
          unsigned char* st = a_smem + (size_t)stage * a_stage_bytes;
          for (int ps = ct; ps < p.pix; ps += 128) {
            int n, ih, iw;
            coords(tile, ps, n, ih, iw);
            float4 v[CH];
            if (n < p.n_img && ih >= 0 && ih < p.h_in && iw >= 0 && iw < p.w_in) {
              const float4* src = reinterpret_cast<const float4*>(p.x + (((size_t)n * p.h_in + ih) * p.w_in + iw) * p.kdim + kc * KC);
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = __ldg(src + j);
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            convert_store(st, ps, v);
          }
          tc::fence_proxy_async();
          tc::mbar_arrive(&a_full[stage]);
          if (++stage == p.sa) { stage = 0; phase ^= 1u; }
        }
      }
    } else {
      auto issue = [&](int tile, int kc, int slot) {
        unsigned char* sg = stg_smem + (size_t)slot * stg_bytes;
        for (int ps = ct; ps < p.pix; ps += 128) {
          int n, ih, iw;
          coords(tile, ps, n, ih, iw);
          const bool ok = n < p.n_img && ih >= 0 && ih < p.h_in && iw >= 0 && iw < p.w_in;
          const float* src = ok ? p.x + (((size_t)n * p.h_in + ih) * p.w_in + iw) * p.kdim + kc * KC : p.x;
#pragma unroll
          for (int j = 0; j < CH; ++j)
            tc::cp_async16(sg + (size_t)ps * (KC * 4) + (size_t)((j ^ swz(ps)) * 16), src + j * 4, ok ? 16u : 0u);
        }
      };
      // prefetch iterator (pt, pk) runs p.ds items ahead of the convert iterator (tile, kc)
      int pt = blockIdx.x, pk = 0, slot_pf = 0;
      for (int i = 0; i < p.ds; ++i) {
        if (pt < p.ntiles) {
          issue(pt, pk, slot_pf);
          if (++pk == nchunks) { pk = 0; pt += gridDim.x; }
        }
        tc::cp_async_commit();
        if (++slot_pf == p.ds) slot_pf = 0;
      }
      int slot_cv = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nchunks; ++kc) {
          // groups are committed one per item, so "all but the newest ds-1" == the item being converted
          if (p.ds == 1) tc::cp_async_wait<0>();
          else if (p.ds == 2) tc::cp_async_wait<1>();
          else tc::cp_async_wait<2>();
          tc::mbar_wait(&a_empty[stage], phase ^ 1u);
          unsigned char* st = a_smem + (size_t)stage * a_stage_bytes;
          const unsigned char* sg = stg_smem + (size_t)slot_cv * stg_bytes;
          for (int ps = ct; ps < p.pix; ps += 128) {
            float4 v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j)
              v[j] = *reinterpret_cast<const float4*>(sg + (size_t)ps * (KC * 4) + (size_t)((j ^ swz(ps)) * 16));
            convert_store(st, ps, v);
          }
          tc::fence_proxy_async();
          tc::mbar_arrive(&a_full[stage]);
          if (++stage == p.sa) { stage = 0; phase ^= 1u; }
          // refill the slot just consumed
          if (pt < p.ntiles) {
            issue(pt, pk, slot_cv);
            if (++pk == nchunks) { pk = 0; pt += gridDim.x; }
          }
          tc::cp_async_commit();
          if (++slot_cv == p.ds) slot_cv = 0;
        }
      }
      tc::cp_async_wait<0>();
    }
  } else if (warp == 8) {
    // ============================== weight blocks: bulk copies per (chunk, tap) =========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const size_t src_plane = (size_t)p.ndim * 16;          // bytes of one [n][8] plane in wprep
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        if (p.b_resident && tile != (int)blockIdx.x) break;     // resident weights: one pass fills every stage
        for (int kc = 0; kc < nchunks; ++kc) {
          for (int tap = 0; tap < 9; ++tap) {
            if (!p.b_resident) tc::mbar_wait(&b_empty[stage], phase ^ 1u);
            tc::mbar_arrive_expect_tx(&b_full[stage], b_stage_bytes);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wprep) +
                                       ((size_t)kc * 9 + tap) * (NSPLIT * Q) * src_plane + (size_t)n0 * 16;
            unsigned char* dst = b_smem + (size_t)stage * b_stage_bytes;
            if (p.nt == p.ndim) {
              tc::bulk_g2s(dst, src, b_stage_bytes, &b_full[stage]);
            } else {
#pragma unroll
              for (int pl = 0; pl < NSPLIT * Q; ++pl) tc::bulk_g2s(dst + (size_t)pl * plane_b, src + (size_t)pl * src_plane, plane_b, &b_full[stage]);
            }
            if (++stage == p.sb) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ============================== MMA issue ==========================================================
    // One thread feeds the tensor core.  With N = 32 an MMA is only ~16 cycles of tensor work, so the
    // issue loop is kept to an add or two per MMA: descriptors are a per-stage base (start address in
    // 16-byte units in the low word) plus tap / K-slice / split offsets.  The WHOLE warp runs this control
    // flow (so the compiler keeps descriptors in uniform registers); one elected lane issues.
    {
      const uint32_t idesc = tc::idesc_bf16_f32(p.nt, 0, 0);
      const uint64_t a_desc0 = tc::smem_desc(tc::smem_u32(a_smem), plane_a, G::GSTRIDE * 16u);
      const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(b_smem), plane_b, 128u);
      const uint32_t a_stage16 = a_stage_bytes >> 4, b_stage16 = b_stage_bytes >> 4;
      const uint32_t plane_a16 = plane_a >> 4, plane_b16 = plane_b >> 4;
      int sa = 0, sb = 0, ab = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      bool b_ready = false;                          // resident weights: waited for once
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
            if (!b_ready) {
              tc::mbar_wait(&b_full[sb], pb);
              tc::tc_fence_after();
            }
            const uint64_t b_base = b_desc0 + (uint64_t)((uint32_t)sb * b_stage16);
            const int acc = G::tap_acc(tap);
            const uint32_t d = d0 + (uint32_t)(acc * p.nt);
            const uint64_t a_tap = a_base + (uint64_t)(uint32_t)p.tap_off[tap];
            // the first tap that touches an accumulator overwrites it on the first channel chunk
            const bool opens = (G::NACC == 1) ? (tap == 0) : (tap == 0 || tap == 1 || tap == 3 || tap == 4);
            if (tc::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                constexpr int NPROD = (NSPLIT == 2) ? 3 : 6;
                constexpr int PA[6] = {0, 0, 1, 0, 2, 1};
                constexpr int PB[6] = {0, 1, 0, 2, 0, 1};
#pragma unroll
                for (int pr = 0; pr < NPROD; ++pr) {
                  const uint64_t da = a_tap + (uint64_t)((uint32_t)(PA[pr] * Q + 2 * ks) * plane_a16);
                  const uint64_t db = b_base + (uint64_t)((uint32_t)(PB[pr] * Q + 2 * ks) * plane_b16);
                  const uint32_t accum = (opens && ks == 0 && pr == 0) ? acc_rest : 1u;
                  tc::mma_bf16(d, da, db, idesc, accum);
                }
              }
              if (!p.b_resident) tc::mma_commit(&b_empty[sb]);
              if (tap == 8) {
                tc::mma_commit(&a_empty[sa]);
                if (kc == nchunks - 1) tc::mma_commit(&acc_full[ab]);
              }
            }
            __syncwarp();
            if (++sb == p.sb) { sb = 0; pb ^= 1u; }
          }
          if (++sa == p.sa) { sa = 0; pa ^= 1u; }
        }
        if (++ab == p.nbuf) { ab = 0; pacc ^= 1u; }
        if (p.b_resident) b_ready = true;
      }
    }
  } else {
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
  if (warp == 9) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
