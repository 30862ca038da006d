// tcgen05 stride-1 3x3 convolution for THIN layers (32 / 64 output channels), "kw-stacked".
//
// conv_tc.cuh issues one MMA pair per (tap, K slice) with N = Cout: at N = 32 the instruction is bound by the
// 4 KB fetch of its A operand from shared memory (49 cycles against 16 of math, profiles/mma_timing_r1.txt), and
// the nine taps re-read the same halo nine times.  Here the three kw taps of a filter row share ONE A fetch:
//
//   tile   = 8 output rows x 14 output columns of one image; staged halo = 10 rows x 16 columns, row pitch
//            exactly 16 pixels, so TMEM lane m = 16 r + c' <-> staged pixel (r + kh, c') and the M operand of tap
//            row kh is the plain window starting kh rows down (start-address shift, group stride 128 bytes);
//   B      = [hi: kw0 kw1 kw2 | lo: kw0 kw1 kw2] x Cout rows per filter row kh, so one MMA of width 6 Cout ("cat",
//            Cout = 32) computes hi x [hi | lo] for all three kw, followed by lo x hi of width 3 Cout
//            (Cout = 64: three MMAs of width 192);
//   acc    P_kw[r][c'] = sum_{kh, cin} x[r + kh - 1, x0 - 1 + c'] W[kh][kw]   (column block kw of the accumulator);
//   out    y[r][x0 + j] = P_0[r][j] + P_1[r][j + 1] + P_2[r][j + 2], j < 14: the epilogue thread of lane (r, j)
//            takes P_1 / P_2 from its neighbours with shfl_down 1 / 2 (a 16-lane row never crosses a warp); two
//            epilogue warpgroups split the channels of every tile (640 threads).
//
// 12 (Cout = 32) or 18 (Cout = 64) MMAs per 32-channel chunk and tile instead of 36 / 54, and 102 KB instead of
// 198 KB of operand fetches per tile.  Data path, rings and warp roles are those of conv_tc.cuh.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

struct TckParams {
  const __nv_bfloat16* wprep;     // [kdim/32][3 kh][4 q][hi | lo][3 kw][nt][8]
  const float* bias;
  int n_img, h, w, kdim, nt;
  float alpha;
  int act;
  int tiles_h, tiles_w, ntiles;
  int sa, sb, ds;                 // ring depths: operand stages, weight stages, raw slots
  int b_resident, cat, tmem_cols;
  int nbuf;                       // accumulator buffers (2 .. 4): nbuf * (cat ? 6 : 3) * nt <= 512 TMEM columns
  // fused epilogues (TcEpi of conv_tc.cuh): MASK reads `aux` (shaped like y) through `tmaux` into a ring of `aux_k`
  // 16 KB slots filled by warp 15; PNF writes rvec
  int epi, aux_k;
  float eps;
  float* rvec;                    // [n, h, w]
#ifdef GS_TC_PROF
  unsigned long long* prof;       // [role][wait0, wait1, wait2, total] cycles
#endif
};

constexpr int TCK_PIX = 160;                       // staged = raw pixels per tile (10 rows x 16 columns)
constexpr int TCK_RAW = 160 * 128;                 // bytes of one raw slot (20480: 1024-aligned)
constexpr int TCK_OUT = 16384;                     // staging tile (8 x 14 pixels x 32 channels fp32 = 14336, padded)
constexpr int TCK_THREADS = 640;                   // conv_tc's 16 warps + a second epilogue warpgroup (warps 16-19)

// Weight pre-pass for the kw-stacked layout: plane index ((kc * 3 + kh) * 4 + q), plane = [split][kw][nt][8].
__device__ __forceinline__ void conv_tck_prep_elem(size_t i, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim,
                                                   int nt, int w_is_kn, int flip) {
  int e = (int)(i % 8);
  size_t r = i / 8;
  int n = (int)(r % nt);
  r /= nt;
  int kw = (int)(r % 3);
  r /= 3;
  int q = (int)(r % 4);
  r /= 4;
  int kh = (int)(r % 3);
  int kc = (int)(r / 3);
  int k = kc * 32 + q * 8 + e;
  int tap = kh * 3 + kw;
  int st = flip ? 8 - tap : tap;
  float v = w_is_kn ? w[((size_t)st * kdim + k) * nt + n] : w[((size_t)st * nt + n) * kdim + k];
  __nv_bfloat16 hi, lo;
  tc::split_bf16(v, hi, lo);
  const size_t plane = (size_t)6 * nt * 8;
  const size_t base = (((size_t)kc * 3 + kh) * 4 + q) * plane;
  out[base + ((size_t)kw * nt + n) * 8 + e] = hi;
  out[base + ((size_t)(3 + kw) * nt + n) * 8 + e] = lo;
}

__device__ __forceinline__ void conv_tck_prep_group(size_t g, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim,
                                                    int nt, int w_is_kn, int flip) {
  size_t r = g;
  int n = (int)(r % nt);
  r /= nt;
  int kw = (int)(r % 3);
  r /= 3;
  int q = (int)(r % 4);
  r /= 4;
  int kh = (int)(r % 3);
  int kc = (int)(r / 3);
  const int k0 = kc * 32 + q * 8;
  const int tap = kh * 3 + kw;
  const int st = flip ? 8 - tap : tap;
  float v[8];
  if (w_is_kn) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = w[((size_t)st * kdim + k0 + e) * nt + n];
  } else {
    const float4* src = reinterpret_cast<const float4*>(w + ((size_t)st * nt + n) * kdim + k0);
    const float4 a = src[0], b = src[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  uint4 h4, l4;
  tc::split2_bf16(v[0], v[1], h4.x, l4.x);
  tc::split2_bf16(v[2], v[3], h4.y, l4.y);
  tc::split2_bf16(v[4], v[5], h4.z, l4.z);
  tc::split2_bf16(v[6], v[7], h4.w, l4.w);
  const size_t plane = (size_t)6 * nt * 8;
  const size_t base = (((size_t)kc * 3 + kh) * 4 + q) * plane;
  *reinterpret_cast<uint4*>(out + base + ((size_t)kw * nt + n) * 8) = h4;
  *reinterpret_cast<uint4*>(out + base + ((size_t)(3 + kw) * nt + n) * 8) = l4;
}

__global__ void conv_tck_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int kdim, int nt,
                                     int w_is_kn, int flip) {
  const size_t total = (size_t)9 * kdim * nt;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    conv_tck_prep_elem(i, w, out, kdim, nt, w_is_kn, flip);
}

// Re-split of up to 48 cached parameter weights in ONE launch (gs_conv_weight_cache_refresh): block b works on the job
// whose block range holds it, 2048 elements per block.
struct PrepJob {
  const float* w;
  __nv_bfloat16* out;
  int kdim, ndim, nt, kc, kn, flip;     // kc: 16 / 32 = conv_tc layout with that chunk, 1032 = the kw-stacked layout
  int block0;                           // first block of this job
};
struct PrepBatch {
  int njobs, nblocks;
  PrepJob jobs[48];
};
constexpr int PREP_ELEMS_PER_BLOCK = 2048;
__global__ void __launch_bounds__(256) conv_prep_batch_kernel(const __grid_constant__ PrepBatch b) {
  int j = 0;
  while (j + 1 < b.njobs && (int)blockIdx.x >= b.jobs[j + 1].block0) ++j;
  const PrepJob& job = b.jobs[j];
  // a thread handles groups of 8 consecutive contraction channels: one set of index divisions, 16-byte stores
  const size_t groups = (size_t)9 * job.kdim * job.ndim / 8;
  const size_t g0 = (size_t)((int)blockIdx.x - job.block0) * (PREP_ELEMS_PER_BLOCK / 8);
#pragma unroll 1
  for (int t = threadIdx.x; t < PREP_ELEMS_PER_BLOCK / 8; t += 256) {
    const size_t g = g0 + t;
    if (g >= groups) break;
    if (job.kc == 1032) conv_tck_prep_group(g, job.w, job.out, job.kdim, job.nt, job.kn, job.flip);
    else if (job.kc == 32) conv_tc_prep_group<32>(g, job.w, job.out, job.kdim, job.ndim, job.nt, job.kn, job.flip);
    else conv_tc_prep_group<16>(g, job.w, job.out, job.kdim, job.ndim, job.nt, job.kn, job.flip);
  }
}

// TPS: filter rows (kh) per weight stage: 3 (one stage per chunk) or 1; EPI: TcEpi (compiled in: the plain epilogue
// keeps its instruction count)
template <int TPS, int CAT, int EPI>
__global__ void __launch_bounds__(TCK_THREADS, 1) conv_tck_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                                                                 const __grid_constant__ CUtensorMap tmaux, const TckParams p) {
  constexpr int KC = 32, Q = 4;
  extern __shared__ unsigned char tck_smem_raw[];
  __shared__ uint64_t raw_full[TC_MAX_STAGES], raw_empty[TC_MAX_STAGES], a_full[TC_MAX_STAGES], a_empty[TC_MAX_STAGES];
  __shared__ uint64_t b_full[TC_MAX_BSTAGES], b_empty[TC_MAX_BSTAGES];
  __shared__ uint64_t acc_full[4], acc_empty[4], aux_full[TC_MAX_AUX], aux_empty[TC_MAX_AUX];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[64];
  __shared__ float ss_x[2][128];          // PNF epilogue: per-pixel partial sums of squares of the two channel halves

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t plane_a = TCK_PIX * 16u;
  const uint32_t plane_b = (uint32_t)(6 * p.nt) * 16u;            // [hi | lo] x 3 kw x nt rows of 16 bytes
  constexpr uint32_t a_stage_bytes = 2 * Q * plane_a;               // [split][q] planes
  const uint32_t b_row_bytes = (uint32_t)Q * plane_b;               // one filter row kh
  const uint32_t b_stage_bytes = (uint32_t)TPS * b_row_bytes;
  unsigned char* smem = tck_smem_raw + ((1024u - (tc::smem_u32(tck_smem_raw) & 1023u)) & 1023u);
  unsigned char* out_smem = smem;                                   // 2 staging tiles
  unsigned char* aux_smem = smem + 2 * TCK_OUT;                     // aux_k slots of TCK_OUT bytes (MASK epilogue)
  unsigned char* raw_smem = aux_smem + (size_t)p.aux_k * TCK_OUT;
  unsigned char* a_smem = raw_smem + (size_t)p.ds * TCK_RAW;
  unsigned char* b_smem = a_smem + (size_t)p.sa * a_stage_bytes;
  const int nchunks = p.kdim / KC;
  constexpr int GROUPS = 3 / TPS;

  if (tid == 0) {
    for (int s = 0; s < p.ds; ++s) { tc::mbar_init(&raw_full[s], 1); tc::mbar_init(&raw_empty[s], TC_CONV_WARPS * 32); }
    for (int s = 0; s < p.sa; ++s) { tc::mbar_init(&a_full[s], TC_CONV_WARPS * 32); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.sb; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 4; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 256); }
    for (int s = 0; s < TC_MAX_AUX; ++s) { tc::mbar_init(&aux_full[s], 1); tc::mbar_init(&aux_empty[s], 256); }
    tc::mbar_fence_init();
  }
  for (int c = tid; c < p.nt; c += TCK_THREADS) bias_s[c] = p.bias ? p.bias[c] : 0.0f;
  if (warp == TC_MMA_WARP0) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp == 12 && lane == 0) {
    tc::prefetch_tmap(&tmx);
    tc::prefetch_tmap(&tmy);
    if (EPI == TC_EPI_MASK) tc::prefetch_tmap(&tmaux);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int acc_cols = (CAT ? 6 : 3) * p.nt;      // columns per accumulator buffer

  if (warp >= 4 && warp < 12) {
    // ============================== fp32 -> bf16 hi/lo split (staged position = raw position) =============
    const int cw = warp - 4;
    const int q = cw & 3;
    const int p_first = (cw >> 2) * 32 + lane;
    int stage = 0, rs = 0;
    uint32_t aph = 0, rph = 0;
    TC_PROF_DECL
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int kc = 0; kc < nchunks; ++kc) {
        TC_WAIT(&raw_full[rs], rph, 0);
        TC_WAIT(&a_empty[stage], aph ^ 1u, 1);
        const unsigned char* raw = raw_smem + (size_t)rs * TCK_RAW;
        unsigned char* st = a_smem + (size_t)stage * a_stage_bytes;
        unsigned char* st_hi = st + (size_t)q * plane_a;
        unsigned char* st_lo = st + (size_t)(Q + q) * plane_a;
        // 160 pixels / 64 per pass: three items per thread (the last one predicated), loads first
        float4 v[3][2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int ps = p_first + 64 * k;
          if (ps < TCK_PIX) {
            const unsigned char* row = raw + (size_t)ps * 128;
            const int sw = ps & 7;
            v[k][0] = *reinterpret_cast<const float4*>(row + (((2 * q) ^ sw) << 4));
            v[k][1] = *reinterpret_cast<const float4*>(row + (((2 * q + 1) ^ sw) << 4));
          }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int ps = p_first + 64 * k;
          if (ps < TCK_PIX) {
            uint4 h4, l4;
            tc::split2_bf16(v[k][0].x, v[k][0].y, h4.x, l4.x);
            tc::split2_bf16(v[k][0].z, v[k][0].w, h4.y, l4.y);
            tc::split2_bf16(v[k][1].x, v[k][1].y, h4.z, l4.z);
            tc::split2_bf16(v[k][1].z, v[k][1].w, h4.w, l4.w);
            *reinterpret_cast<uint4*>(st_hi + (size_t)ps * 16) = h4;
            *reinterpret_cast<uint4*>(st_lo + (size_t)ps * 16) = l4;
          }
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&a_full[stage]);
        tc::mbar_arrive(&raw_empty[rs]);
        if (++stage == p.sa) { stage = 0; aph ^= 1u; }
        if (++rs == p.ds) { rs = 0; rph ^= 1u; }
      }
    }
    TC_PROF_FLUSH(p.prof, 0, warp == 4 && lane == 0);
  } else if (warp == 12) {
    // ============================== halo tiles: one TMA box (32 ch, 16, 1, 10) per (tile, chunk) =========
    if (lane == 0) {
      int rs = 0;
      uint32_t rph = 0;
      TC_PROF_DECL
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int t = tile;
        const int tw_ = t % p.tiles_w;
        t /= p.tiles_w;
        const int th_ = t % p.tiles_h;
        const int n = t / p.tiles_h;
        for (int kc = 0; kc < nchunks; ++kc) {
          TC_WAIT(&raw_empty[rs], rph ^ 1u, 0);
          tc::mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)TCK_RAW);
          tc::tma_load_4d(raw_smem + (size_t)rs * TCK_RAW, &tmx, kc * KC, tw_ * 14 - 1, n, th_ * 8 - 1, &raw_full[rs]);
          if (++rs == p.ds) { rs = 0; rph ^= 1u; }
        }
      }
      TC_PROF_FLUSH(p.prof, 1, true);
    }
  } else if (warp == 13) {
    // ============================== weight blocks ======================================================
    if (lane == 0) {
      const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wprep);
      if (p.b_resident) {
        for (int i = 0; i < nchunks * GROUPS; ++i) {
          tc::mbar_arrive_expect_tx(&b_full[i], b_stage_bytes);
          tc::bulk_g2s(b_smem + (size_t)i * b_stage_bytes, wsrc + (size_t)i * b_stage_bytes, b_stage_bytes, &b_full[i]);
        }
      } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
          for (int i = 0; i < nchunks * GROUPS; ++i) {
            tc::mbar_wait(&b_empty[stage], phase ^ 1u);
            tc::mbar_arrive_expect_tx(&b_full[stage], b_stage_bytes);
            tc::bulk_g2s(b_smem + (size_t)stage * b_stage_bytes, wsrc + (size_t)i * b_stage_bytes, b_stage_bytes, &b_full[stage]);
            if (++stage == p.sb) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == TC_MMA_WARP0) {
    // ============================== MMA issue ==========================================================
    const uint32_t idesc_w = tc::idesc_bf16_f32((CAT ? 6 : 3) * p.nt, 0, 0);   // hi x [hi | lo] x 3 kw  /  x 3 kw
    const uint32_t idesc_n = tc::idesc_bf16_f32(3 * p.nt, 0, 0);
    const uint64_t a_desc0 = tc::smem_desc(tc::smem_u32(a_smem), plane_a, 128u);
    const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(b_smem), plane_b, 128u);
    constexpr uint32_t a_stage16 = a_stage_bytes >> 4, plane_a16 = plane_a >> 4;
    const uint32_t b_stage16 = b_stage_bytes >> 4, b_row16 = b_row_bytes >> 4, plane_b16 = plane_b >> 4;
    const uint32_t b_lo16 = (uint32_t)(3 * p.nt);                               // lo rows follow the 3 nt hi rows
    constexpr uint32_t a_lo16 = Q * plane_a16;
    int sa = 0, sb = 0, ab = 0;
    uint32_t pa = 0, pb = 0, pacc = 0;
    bool b_ready = false;
    TC_PROF_DECL
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      TC_WAIT(&acc_empty[ab], pacc ^ 1u, 0);
      tc::tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(ab * acc_cols);
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
              const int kh = g * TPS + tt;
              const uint64_t a_row = a_base + (uint64_t)(uint32_t)(kh * 16);          // kh staged rows of 256 bytes
              const uint64_t b_row = b_base + (uint64_t)((uint32_t)tt * b_row16);
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t a_hi = a_row + (uint64_t)((uint32_t)(2 * ks) * plane_a16);
                const uint64_t b_k = b_row + (uint64_t)((uint32_t)(2 * ks) * plane_b16);
                const uint32_t accum = (kh == 0 && ks == 0) ? acc_rest : 1u;
                if (CAT) {
                  tc::mma_bf16(d, a_hi, b_k, idesc_w, accum);
                  tc::mma_bf16(d, a_hi + a_lo16, b_k, idesc_n, 1u);
                } else {
                  tc::mma_bf16(d, a_hi, b_k, idesc_n, accum);
                  tc::mma_bf16(d, a_hi, b_k + b_lo16, idesc_n, 1u);
                  tc::mma_bf16(d, a_hi + a_lo16, b_k, idesc_n, 1u);
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
          if (++sb == p.sb) { sb = 0; pb ^= 1u; }
        }
        if (++sa == p.sa) { sa = 0; pa ^= 1u; }
      }
      if (++ab == p.nbuf) { ab = 0; pacc ^= 1u; }
      if (p.b_resident) b_ready = true;
    }
    TC_PROF_FLUSH(p.prof, 2, lane == 0);
  } else if (warp == 15) {
    // ============================== aux tiles of the MASK epilogue: one TMA box per (tile, chunk) ==========
    if (EPI == TC_EPI_MASK && lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int t = tile;
        const int tw_ = t % p.tiles_w;
        t /= p.tiles_w;
        const int th_ = t % p.tiles_h;
        const int n = t / p.tiles_h;
        for (int c0 = 0; c0 < p.nt; c0 += 32) {
          tc::mbar_wait(&aux_empty[slot], ph ^ 1u);
          tc::mbar_arrive_expect_tx(&aux_full[slot], 14336u);
          tc::tma_load_4d(aux_smem + (size_t)slot * TCK_OUT, &tmaux, c0, tw_ * 14, n, th_ * 8, &aux_full[slot]);
          if (++slot == p.aux_k) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp < 4 || warp >= 16) {
    // ============================== epilogue: TMEM -> shift-add over kw -> alpha, bias, leaky-relu -> TMA store =====
    // Both warpgroups (warps 0-3 and 16-19) work on EVERY tile: a warp reads the TMEM lane quarter warp % 4, so the two
    // warps of a quarter split the COLUMNS -- group g owns channels [16 g, 16 g + 16) of each 32-channel chunk.  Per
    // (tile, kw) the hi / lo column blocks are fetched with one TMEM round trip, and the accumulator buffer goes back
    // to the MMA warp as soon as it is in registers, before the arithmetic and the store (measured with the stage
    // profile, tools/tc_stage_profile.py: the buffer cycle MMA -> drain was the pace of this kernel, not the MMAs).
    const int grp = warp >= 16 ? 1 : 0;
    const int wq = warp & 3;                                         // TMEM lane quarter of this warp
    const int m = wq * 32 + lane;
    const int r = m >> 4, cp = m & 15;
    const bool writer = cp < 14;
    const bool issuer = (grp == 0 && wq == 0 && lane == 0);
    const int srow = r * 14 + cp;                                    // row of the dense [8][14] staging tile
    const int sw = srow & 7;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const int nchunks_out = p.nt >> 5;                               // 1 or 2
    const float inv_nt = 1.0f / (float)p.nt;
    int ab = 0;
    uint32_t pacc = 0, seq = 0;                                      // seq: 32-channel chunk number in this CTA's sequence
    TC_PROF_DECL
#ifdef GS_TC_PROF
    long long pe0_ = 0, pe1_ = 0, pe2_ = 0;
#endif
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      const int n = t / p.tiles_h;
      TC_WAIT(&acc_full[ab], pacc, 0);
      tc::tc_fence_after();
      const uint32_t acc0 = tmem_base + lane_base + (uint32_t)(ab * acc_cols + 16 * grp);
      float v[2][16];
#ifdef GS_TC_PROF
      const long long te0_ = clock64();
#endif
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nchunks_out) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            uint32_t pk[16], p2[16];
            tc::tmem_ld16_issue(acc0 + (uint32_t)(kw * p.nt + 32 * c), pk);
            if (CAT) {
              tc::tmem_ld16_issue(acc0 + (uint32_t)((3 + kw) * p.nt + 32 * c), p2);
              tc::tmem_ld_wait(pk, p2);
            } else {
              tc::tmem_ld_wait(pk);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = __uint_as_float(pk[j]);
              if (CAT) x += __uint_as_float(p2[j]);
              if (kw == 0) v[c][j] = x;
              else v[c][j] += __shfl_down_sync(0xffffffffu, x, kw);
            }
          }
        }
      }
#ifdef GS_TC_PROF
      const long long te1_ = clock64();
      pe0_ += te1_ - te0_;
#endif
      // the accumulator is in registers: the MMA warp may start the tile after next
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[ab]);
      if (++ab == p.nbuf) { ab = 0; pacc ^= 1u; }
#ifdef GS_TC_PROF
      pe1_ += clock64() - te1_;
#endif

      float rscale = 1.0f;
      if (EPI == TC_EPI_PNF) {
        // mean square of the activated outputs of this pixel over ALL channels: the two groups exchange partial sums
        float ss = 0.0f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < nchunks_out) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float o = gs_lrelu(fmaf(v[c][j], p.alpha, bias_s[32 * c + 16 * grp + j]));
              ss = fmaf(o, o, ss);
            }
          }
        }
        ss_x[grp][m] = ss;
        asm volatile("bar.sync 3, 256;" ::: "memory");
        ss += ss_x[grp ^ 1][m];                   // rewritten only after this tile's staging barrier
        rscale = 1.0f / sqrtf(ss * inv_nt + p.eps);
        const int px = tw_ * 14 + cp;
        if (grp == 0 && writer && px < p.w) p.rvec[((size_t)n * p.h + th_ * 8 + r) * p.w + px] = rscale;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nchunks_out) {
          const uint32_t buf = seq & 1u;
          unsigned char* row = out_smem + (size_t)buf * TCK_OUT + (size_t)srow * 128;
          if (EPI == TC_EPI_MASK) {
            const uint32_t aslot = seq % (uint32_t)p.aux_k, aph = (seq / (uint32_t)p.aux_k) & 1u;
            TC_WAIT(&aux_full[aslot], aph, 2);
            if (writer) {
              const unsigned char* arow = aux_smem + (size_t)aslot * TCK_OUT + (size_t)srow * 128;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const int ch16 = 4 * grp + (j >> 2);
                const float4 ax = *reinterpret_cast<const float4*>(arow + ((ch16 ^ sw) << 4));
                float4 o;
                o.x = v[c][j + 0] * p.alpha * gs_lrelu_slope(ax.x); o.y = v[c][j + 1] * p.alpha * gs_lrelu_slope(ax.y);
                o.z = v[c][j + 2] * p.alpha * gs_lrelu_slope(ax.z); o.w = v[c][j + 3] * p.alpha * gs_lrelu_slope(ax.w);
                *reinterpret_cast<float4*>(row + ((ch16 ^ sw) << 4)) = o;
              }
            }
            tc::mbar_arrive(&aux_empty[aslot]);
          } else if (writer) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int ch16 = 4 * grp + (j >> 2);
              const float4 bv = *reinterpret_cast<const float4*>(&bias_s[32 * c + 16 * grp + j]);
              float4 o;
              o.x = fmaf(v[c][j + 0], p.alpha, bv.x); o.y = fmaf(v[c][j + 1], p.alpha, bv.y);
              o.z = fmaf(v[c][j + 2], p.alpha, bv.z); o.w = fmaf(v[c][j + 3], p.alpha, bv.w);
              if (p.act == 1) { o.x = gs_lrelu(o.x); o.y = gs_lrelu(o.y); o.z = gs_lrelu(o.z); o.w = gs_lrelu(o.w); }
              if (EPI == TC_EPI_PNF) { o.x *= rscale; o.y *= rscale; o.z *= rscale; o.w *= rscale; }
              *reinterpret_cast<float4*>(row + ((ch16 ^ sw) << 4)) = o;
            }
          }
#ifdef GS_TC_PROF
          const long long te2_ = clock64();
#endif
          tc::fence_proxy_async();
#ifdef GS_TC_PROF
          pe2_ += clock64() - te2_;
#endif
          // Staging tile seq & 1 was last read by the store of chunk seq - 2; the issuer has waited for that store
          // before the barrier of chunk seq - 1.  Here it waits for the store of chunk seq - 1, whose tile chunk
          // seq + 1 rewrites after this barrier.
          TC_PROF_BEGIN(1)
          if (issuer) tc::bulk_wait_read<0>();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          TC_PROF_END(1)
          if (issuer) {
            tc::tma_store_4d(&tmy, out_smem + (size_t)buf * TCK_OUT, 32 * c, tw_ * 14, n, th_ * 8);
            tc::bulk_commit();
          }
          ++seq;
        }
      }
    }
    TC_PROF_FLUSH(p.prof, 3 + grp, wq == 0 && lane == 0);
#ifdef GS_TC_PROF
    if (p.prof != nullptr && grp == 0 && wq == 0 && lane == 0) {
      atomicAdd(p.prof + 5 * 4 + 0, (unsigned long long)pe0_);                   // TMEM loads + shift-add
      atomicAdd(p.prof + 5 * 4 + 1, (unsigned long long)pe1_);                   // fence + arrive on acc_empty
      atomicAdd(p.prof + 5 * 4 + 2, (unsigned long long)pe2_);                   // fence.proxy.async
      atomicAdd(p.prof + 5 * 4 + 3, (unsigned long long)(clock64() - pt0_));
    }
#endif
    if (issuer) tc::bulk_wait<0>();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP0) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
