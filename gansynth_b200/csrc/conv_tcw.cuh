// tcgen05 / TMEM filter-gradient ("W form") of the 3x3 convolutions, bf16x3, for sm_100a.
//
//   dw(kh, kw, a, b) = alpha * sum_{n,oy,ox} big[n, oy*S + kh - pb, ox*S + kw - pb, a] * small[n, oy, ox, b]
//
// Per tap this is a GEMM whose contraction index is the PIXEL.  Both operands sit in shared memory as
// 16-byte channel vectors (8 channels of one pixel), 8 consecutive pixels of an image row contiguous: for a
// GEMM contracted over pixels that is the MN-major SWIZZLE_NONE core-matrix layout (channel chunks along
// M / N, 8-pixel groups along K).
//
//   big   staged as [row][q][PW pixels]   (q = 8-channel chunk).  Consecutive (row, q) blocks are equally
//         spaced, so the M index of ONE instruction can run over (kh, channel): the three kh taps are three
//         runs of Q chunks one staged row apart -- "kh-stacked", no copies, M = 3 * channels (<= 128 rows
//         per job).  kw is a start-address shift (stride 2: even / odd column planes).
//   small staged as [row][hi q.. | lo q..][8 pixels]: hi and lo chunks adjacent, so hi x [hi | lo] is one
//         MMA of width 2 * NB ("cat", as in conv_tc.cuh) when the accumulators fit TMEM.
//
// "N-stacked" variant (stride 1, NB = 32): the three kw taps move to the N side.  `small` is staged as three
// column-shifted copies [row][hi: kw0 kw1 kw2 | lo: kw0 kw1 kw2][q][8 pixels] (raw box 10 pixels wide), `big` needs no
// column halo, and ONE MMA of width 6 NB (cat) + one of width 3 NB per K step replace the three kw groups (measured:
// 112 -> 104 us on the top layer; no gain at NB = 64, and 12 converter warps instead of 8 buy nothing either).
//
// A job (blockIdx.y) = (kh range, big-channel block of <= 128, small-channel block NB <= 128): one 128-row
// accumulator per kw (3 * NB * (1 + cat) <= 512 TMEM columns).  grid.x splits the pixel tiles; every CTA
// accumulates over its whole pixel range and adds its partial filter gradient to dw with (vector) fp32
// atomics at the end.
//
// Data path (as in conv_tc.cuh): TMA tiled loads of fp32 boxes (128B swizzle, zero fill outside the image)
// into a raw ring -> 8 converter warps split fp32 -> bf16 hi/lo into the operand ring -> one elected thread
// issues the MMAs -> warps 0-3 drain TMEM at the end.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

struct TcwMaps {
  CUtensorMap big[2];     // box rows differ with the number of kh taps of the job
  CUtensorMap small;
};

struct TcwParams {
  float* dw;
  int n_img, bh, bw, sh, sw, adim, bdim, stride;
  int tpr;                        // tile rows (small side); a tile is tpr x 8 pixels of one image
  int tiles_h, tiles_w, ntiles;
  int mjobs, njobs_n;             // grid.y = mjobs * njobs_n
  int kh0[8], nkh[8], ch0[8], map_id[8];
  int nch, nb;                    // big / small channels per job (multiples of 32, <= 128)
  int cat;
  int nstack;                     // kw taps stacked along N (three shifted copies of `small`)
  int pw, bwraw;                  // staged / raw pixels per big row: 10 / 10 (stride 1), 18 / 17 (stride 2)
  int hr_max;                     // staged big rows of the job with the most kh taps
  int stages, ds, out_ab, tmem_cols;
  float alpha;
  uint32_t raw_big_chunk, raw_small_chunk, raw_slot_bytes;   // raw ring: 32-channel chunks, 1024-aligned
  uint32_t big_lo_off, small_off, stage_bytes;                // operand stage layout (bytes)
  // fused bias gradient: dbias[c] += sum over all pixels of one operand (the layer's pre-activation gradient), taken
  // from the converter warps' registers.  bias_side 1 = `small`, 2 = `big`.  Every pixel must be counted once: the
  // small side by the jobs with mj == 0 (raw columns 1..8 in the N-stacked form, whose boxes overlap), the big side
  // by the jobs with nj == 0 on the raw rows [bias_r0, bias_r1) of job mj (the kh jobs of a channel block overlap)
  // and the raw columns [bias_c0, bias_c1).
  float* dbias;
  int bias_side, bias_c0, bias_c1;
  int bias_r0[8], bias_r1[8];
#ifdef GS_TC_PROF
  unsigned long long* prof;       // [role][wait0, wait1, wait2, total] cycles
#endif
};

#ifndef TCW_CW
#define TCW_CW 8                     // converter warps (a build-time experiment knob: -DTCW_CW=16)
#endif
constexpr int TCW_CONV_WARPS = TCW_CW;
constexpr int TCW_THREADS = 32 * (TCW_CONV_WARPS + 2);        // + the TMA warp and the MMA warp
constexpr int TCW_CONV_THREADS = 32 * TCW_CONV_WARPS;
constexpr int TCW_MAX_STAGES = 4;

__global__ void __launch_bounds__(TCW_THREADS, 1) conv_tcw_kernel(const __grid_constant__ TcwMaps maps, const TcwParams p) {
  extern __shared__ unsigned char tcw_smem_raw[];
  __shared__ uint64_t raw_full[TCW_MAX_STAGES], raw_empty[TCW_MAX_STAGES], full[TCW_MAX_STAGES], empty[TCW_MAX_STAGES], done;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint16_t big_tab[640];       // raw big pixel -> staged position (16-byte units, q = 0); bit 15: counts for dbias
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // job decode
  const int mj = blockIdx.y / p.njobs_n, nj = blockIdx.y % p.njobs_n;
  const int kh0 = p.kh0[mj], nkh = p.nkh[mj], ch0 = p.ch0[mj];
  const int nb0 = nj * p.nb;
  const int S = p.stride;
  const int qj = p.nch >> 3, qb = p.nb >> 3;
  const int hr_rows = S * (p.tpr - 1) + nkh;
  const int swraw = p.nstack ? 10 : 8;                       // raw small box width
  const int nbig_px = hr_rows * p.bwraw, nsmall_px = p.tpr * swraw;
  const int big_chunks = p.nch >> 5, small_chunks = p.nb >> 5;
  unsigned char* smem = tcw_smem_raw + ((1024u - (tc::smem_u32(tcw_smem_raw) & 1023u)) & 1023u);
  unsigned char* raw_smem = smem;
  unsigned char* st_smem = smem + (size_t)p.ds * p.raw_slot_bytes;

  if (tid == 0) {
    for (int s = 0; s < p.ds; ++s) { tc::mbar_init(&raw_full[s], 1); tc::mbar_init(&raw_empty[s], TCW_CONV_THREADS); }
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full[s], TCW_CONV_THREADS); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&done, 1);
    tc::mbar_fence_init();
  }
  for (int px = tid; px < nbig_px; px += TCW_THREADS) {
    const int hr = px / p.bwraw, hc = px % p.bwraw;
    const bool counts = p.bias_side == 2 && nj == 0 && hr >= p.bias_r0[mj] && hr < p.bias_r1[mj] && hc >= p.bias_c0 && hc < p.bias_c1;
    big_tab[px] = (uint16_t)((hr * qj * p.pw + (S == 1 ? hc : (hc & 1) * 9 + (hc >> 1))) | (counts ? 0x8000 : 0));
  }
  if (warp == TCW_CONV_WARPS + 1) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp == TCW_CONV_WARPS && lane == 0) {
    tc::prefetch_tmap(&maps.big[p.map_id[mj]]);
    tc::prefetch_tmap(&maps.small);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int acc_w = p.cat ? 2 * p.nb : p.nb;
  const int nchunk_row = p.nstack ? 6 * qb : 2 * qb;         // 16-byte x 8-pixel chunks per staged row of small

  if (warp < TCW_CONV_WARPS) {
    // ============================== fp32 -> bf16 hi/lo split ===========================================
    // item = (pixel, 8-channel chunk q).  A warp works on one q at a time with consecutive pixels on
    // consecutive lanes: conflict-free reads of the swizzled raw rows, contiguous 16-byte stores.
    constexpr int CW = TCW_CONV_WARPS;
    const int wq_b = qj >= CW ? 1 : CW / qj, wq_s = qb >= CW ? 1 : CW / qb;     // warps per chunk (idle warps when CW % q != 0)
    int stage = 0, rs = 0;
    uint32_t ph = 0, rph = 0;
    // fused bias gradient: a warp visits at most two 8-channel chunks of the operand (q, q + 8)
    const bool bias_big = p.bias_side == 2 && nj == 0 && p.bias_r1[mj] > p.bias_r0[mj];
    const bool bias_small = p.bias_side == 1 && mj == 0;
    float bsum[2][8];
    auto bias_add = [](float (&b)[8], const float4& a0, const float4& a1) {
      b[0] += a0.x; b[1] += a0.y; b[2] += a0.z; b[3] += a0.w;
      b[4] += a1.x; b[5] += a1.y; b[6] += a1.z; b[7] += a1.w;
    };
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) bsum[i][j] = 0.0f;
    TC_PROF_DECL
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      TC_WAIT(&raw_full[rs], rph, 0);
      TC_WAIT(&empty[stage], ph ^ 1u, 1);
      const unsigned char* raw = raw_smem + (size_t)rs * p.raw_slot_bytes;
      unsigned char* st = st_smem + (size_t)stage * p.stage_bytes;
      // ---- big: staged [row][q][pw] (hi block, then lo block)
      for (int q = (qj >= CW ? warp : warp / wq_b); q < qj; q += (qj >= CW ? CW : qj)) {
        const unsigned char* chunk = raw + (size_t)(q >> 2) * p.raw_big_chunk;
        const int j0 = 2 * (q & 3);
        for (int px = (qj >= CW ? 0 : (warp % wq_b) * 32) + lane; px < nbig_px; px += 32 * wq_b) {
          const unsigned char* row = chunk + (size_t)px * 128;
          const int sw = px & 7;
          const float4 v0 = *reinterpret_cast<const float4*>(row + ((j0 ^ sw) << 4));
          const float4 v1 = *reinterpret_cast<const float4*>(row + (((j0 + 1) ^ sw) << 4));
          uint4 h4, l4;
          tc::split2_bf16(v0.x, v0.y, h4.x, l4.x);
          tc::split2_bf16(v0.z, v0.w, h4.y, l4.y);
          tc::split2_bf16(v1.x, v1.y, h4.z, l4.z);
          tc::split2_bf16(v1.z, v1.w, h4.w, l4.w);
          const uint32_t tab = big_tab[px];
          const uint32_t d = ((tab & 0x7fffu) + (uint32_t)(q * p.pw)) << 4;
          *reinterpret_cast<uint4*>(st + d) = h4;
          *reinterpret_cast<uint4*>(st + p.big_lo_off + d) = l4;
          if (bias_big && (tab & 0x8000u)) {
            if (q >= CW) bias_add(bsum[1], v0, v1);    // constant indices: the sums stay in registers
            else bias_add(bsum[0], v0, v1);
          }
        }
      }
      // ---- small: staged [row][hi q.. | lo q..][8 pixels]
      unsigned char* ss = st + p.small_off;
      for (int q = (qb >= CW ? warp : warp / wq_s); q < qb; q += (qb >= CW ? CW : qb)) {
        const unsigned char* chunk = raw + (size_t)big_chunks * p.raw_big_chunk + (size_t)(q >> 2) * p.raw_small_chunk;
        const int j0 = 2 * (q & 3);
        for (int px = (qb >= CW ? 0 : (warp % wq_s) * 32) + lane; px < nsmall_px; px += 32 * wq_s) {
          const unsigned char* row = chunk + (size_t)px * 128;
          const int sw = px & 7;
          const float4 v0 = *reinterpret_cast<const float4*>(row + ((j0 ^ sw) << 4));
          const float4 v1 = *reinterpret_cast<const float4*>(row + (((j0 + 1) ^ sw) << 4));
          uint4 h4, l4;
          tc::split2_bf16(v0.x, v0.y, h4.x, l4.x);
          tc::split2_bf16(v0.z, v0.w, h4.y, l4.y);
          tc::split2_bf16(v1.x, v1.y, h4.z, l4.z);
          tc::split2_bf16(v1.z, v1.w, h4.w, l4.w);
          if (bias_small) {
            bool counts = true;
            if (p.nstack) { const int hc = px % 10; counts = hc >= 1 && hc <= 8; }
            if (counts) {
              if (q >= CW) bias_add(bsum[1], v0, v1);
              else bias_add(bsum[0], v0, v1);
            }
          }
          if (!p.nstack) {
            const uint32_t r = (uint32_t)px >> 3, c = (uint32_t)px & 7u;
            const uint32_t d = ((r * 2u * (uint32_t)qb + (uint32_t)q) * 8u + c) << 4;
            *reinterpret_cast<uint4*>(ss + d) = h4;
            *reinterpret_cast<uint4*>(ss + d + ((uint32_t)qb << 7)) = l4;
          } else {
            // raw column hc (image column ox0 - 1 + hc) is pixel c = hc - 2 + kw of the copy for tap kw
            const int r = px / 10, hc = px - r * 10;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int c = hc - 2 + kw;
              if (c >= 0 && c < 8) {
                const uint32_t d = (((uint32_t)(r * nchunk_row + kw * qb + q)) * 8u + (uint32_t)c) << 4;
                *reinterpret_cast<uint4*>(ss + d) = h4;
                *reinterpret_cast<uint4*>(ss + d + ((uint32_t)(3 * qb) << 7)) = l4;
              }
            }
          }
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[stage]);
      tc::mbar_arrive(&raw_empty[rs]);
      if (++stage == p.stages) { stage = 0; ph ^= 1u; }
      if (++rs == p.ds) { rs = 0; rph ^= 1u; }
    }

    TC_PROF_FLUSH(p.prof, 8, warp == 0 && lane == 0);
    // ============================== fused bias gradient: lanes -> one atomic per (warp, channel) =========
    if (bias_big || bias_small) {
      const int qn = bias_big ? qj : qb;
      const int cbase = bias_big ? ch0 : nb0;
      const int q0 = qn >= CW ? warp : warp / (CW / qn);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = q0 + CW * i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = bsum[i][j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && q < qn && (i == 0 || qn > CW)) atomicAdd(p.dbias + cbase + q * 8 + j, v);
        }
      }
    }

    // ============================== drain: TMEM -> atomics into dw =====================================
    if (warp < 4) {
#ifdef GS_TC_PROF
      const long long td0_ = clock64();
#endif
      tc::mbar_wait(&done, 0);
      tc::tc_fence_after();
#ifdef GS_TC_PROF
      const long long td1_ = clock64();
#endif
      const int m = warp * 32 + lane;
      const int g = m >> 3;
      const bool row_ok = g < nkh * qj;
      const int kh = kh0 + g / qj;
      const int a = ch0 + (g % qj) * 8 + (m & 7);
      const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
      for (int kw = 0; kw < 3; ++kw) {
        const int tap = kh * 3 + kw;
#pragma unroll 1
        for (int c0 = 0; c0 < p.nb; c0 += 32) {
          float v[32];
          const int col_hh = p.nstack ? kw * p.nb : kw * acc_w;
          const int col_hl = p.nstack ? (3 + kw) * p.nb : kw * acc_w + p.nb;
          tc::tmem_ld32(tmem_base + lane_base + (uint32_t)(col_hh + c0), v);
          if (p.cat) {
            float v2[32];
            tc::tmem_ld32(tmem_base + lane_base + (uint32_t)(col_hl + c0), v2);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += v2[j];
          }
          if (row_ok) {
            if (p.out_ab) {
              float* dst = p.dw + ((size_t)tap * p.adim + a) * p.bdim + nb0 + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                atomicAdd(reinterpret_cast<float4*>(dst + j),
                          make_float4(v[j] * p.alpha, v[j + 1] * p.alpha, v[j + 2] * p.alpha, v[j + 3] * p.alpha));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                atomicAdd(p.dw + ((size_t)tap * p.bdim + nb0 + c0 + j) * p.adim + a, v[j] * p.alpha);
            }
          }
        }
      }
#ifdef GS_TC_PROF
      if (p.prof != nullptr && tid == 0) {
        atomicAdd(p.prof + 11 * 4 + 0, (unsigned long long)(td1_ - td0_));        // wait for the last MMA
        atomicAdd(p.prof + 11 * 4 + 3, (unsigned long long)(clock64() - td1_));   // TMEM -> atomics
      }
#endif
    }
  } else if (warp == TCW_CONV_WARPS) {
    // ============================== TMA: one box per 32-channel chunk of each operand ====================
    if (lane == 0) {
      int rs = 0;
      uint32_t rph = 0;
      const uint32_t bytes = (uint32_t)(big_chunks * nbig_px + small_chunks * nsmall_px) * 128u;
      const CUtensorMap* mb = &maps.big[p.map_id[mj]];
      TC_PROF_DECL
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int t = tile;
        const int tw_ = t % p.tiles_w;
        t /= p.tiles_w;
        const int th_ = t % p.tiles_h;
        const int n = t / p.tiles_h;
        const int ox0 = tw_ * 8, oy0 = th_ * p.tpr;
        const int bx0 = (S == 1) ? (p.nstack ? ox0 : ox0 - 1) : 2 * ox0;
        const int by0 = (S == 1) ? oy0 - 1 + kh0 : 2 * oy0 + kh0;
        TC_WAIT(&raw_empty[rs], rph ^ 1u, 0);
        tc::mbar_arrive_expect_tx(&raw_full[rs], bytes);
        unsigned char* slot = raw_smem + (size_t)rs * p.raw_slot_bytes;
        for (int c = 0; c < big_chunks; ++c)
          tc::tma_load_4d(slot + (size_t)c * p.raw_big_chunk, mb, ch0 + 32 * c, bx0, n, by0, &raw_full[rs]);
        for (int c = 0; c < small_chunks; ++c)
          tc::tma_load_4d(slot + (size_t)big_chunks * p.raw_big_chunk + (size_t)c * p.raw_small_chunk, &maps.small, nb0 + 32 * c,
                          p.nstack ? ox0 - 1 : ox0, n, oy0, &raw_full[rs]);
        if (++rs == p.ds) { rs = 0; rph ^= 1u; }
      }
      TC_PROF_FLUSH(p.prof, 9, true);
    }
  } else {
    // ============================== MMA issue ==========================================================
    // A (big, M side): chunk stride = pw pixels, 8-pixel K group stride = S staged rows; B (small, N side):
    // chunk stride = 8 pixels, K group stride = one staged row.  One K step = two tile rows.
    const uint32_t n_wide = p.nstack ? (uint32_t)((p.cat ? 6 : 3) * p.nb) : (uint32_t)acc_w;
    const uint32_t n_narrow = p.nstack ? (uint32_t)(3 * p.nb) : (uint32_t)p.nb;
    const uint32_t idesc_w = tc::idesc_bf16_f32((int)n_wide, 1, 1);
    const uint32_t idesc_n = tc::idesc_bf16_f32((int)n_narrow, 1, 1);
    const uint32_t lbo_m = (uint32_t)(S * qj * p.pw) * 16u, sbo_m = (uint32_t)p.pw * 16u;
    const uint32_t lbo_n = (uint32_t)nchunk_row * 128u, sbo_n = 128u;
    const uint64_t m_desc0 = tc::smem_desc(tc::smem_u32(st_smem), lbo_m, sbo_m);
    const uint64_t n_desc0 = tc::smem_desc(tc::smem_u32(st_smem) + p.small_off, lbo_n, sbo_n);
    const uint32_t stage16 = p.stage_bytes >> 4;
    const uint32_t m_lo16 = p.big_lo_off >> 4, n_lo16 = ((uint32_t)(p.nstack ? 3 * qb : qb) << 7) >> 4;
    const uint32_t m_step16 = (2u * lbo_m) >> 4, n_step16 = (2u * lbo_n) >> 4;
    const int ksteps = p.tpr / 2;
    const int ngroups = p.nstack ? 1 : 3;             // kw groups per K step (stacked: all three in one instruction)
    int stage = 0;
    uint32_t ph = 0, accum_first = 0;
    TC_PROF_DECL
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      TC_WAIT(&full[stage], ph, 0);
      tc::tc_fence_after();
      const uint64_t m_base = m_desc0 + (uint64_t)((uint32_t)stage * stage16);
      const uint64_t n_base = n_desc0 + (uint64_t)((uint32_t)stage * stage16);
      if (tc::elect_one()) {
        for (int kw = 0; kw < ngroups; ++kw) {
          const uint32_t kw16 = p.nstack ? 0u : (S == 1) ? (uint32_t)kw : (uint32_t)((kw & 1) * 9 + (kw >> 1));
          uint64_t a_hi = m_base + (uint64_t)kw16;
          uint64_t b_hi = n_base;
          const uint32_t d = tmem_base + (uint32_t)(kw * acc_w);
          uint32_t accum = accum_first;
          if (p.cat) {
#pragma unroll 2
            for (int j = 0; j < ksteps; ++j) {
              tc::mma_bf16(d, a_hi, b_hi, idesc_w, accum);
              tc::mma_bf16(d, a_hi + m_lo16, b_hi, idesc_n, 1u);
              accum = 1u;
              a_hi += m_step16;
              b_hi += n_step16;
            }
          } else {
#pragma unroll 2
            for (int j = 0; j < ksteps; ++j) {
              tc::mma_bf16(d, a_hi, b_hi, idesc_n, accum);
              tc::mma_bf16(d, a_hi, b_hi + n_lo16, idesc_n, 1u);
              tc::mma_bf16(d, a_hi + m_lo16, b_hi, idesc_n, 1u);
              accum = 1u;
              a_hi += m_step16;
              b_hi += n_step16;
            }
          }
        }
        tc::mma_commit(&empty[stage]);
      }
      __syncwarp();
      accum_first = 1u;
      if (++stage == p.stages) { stage = 0; ph ^= 1u; }
    }
    if (tc::elect_one()) tc::mma_commit(&done);
    __syncwarp();
    TC_PROF_FLUSH(p.prof, 10, lane == 0);
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == TCW_CONV_WARPS + 1) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
