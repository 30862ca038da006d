// tcgen05 / TMEM filter-gradient ("W form") of the 3x3 convolutions, bf16x3, for sm_100a.
//
//   dw(tap, a, b) = alpha * sum_{n,oy,ox} big[n, oy*S+kh-pb, ox*S+kw-pb, a] * small[n, oy, ox, b]
//
// Per tap this is a GEMM whose contraction index is the PIXEL.  Both operands sit in shared memory as
// 16-byte channel vectors in pixel-major planes, which for this GEMM is the MN-major SWIZZLE_NONE
// core-matrix layout: channel chunks along M / N (stride = plane), pixels along K (8 consecutive pixels =
// one image-row segment, next segment = next tile row).
//
// Two job layouts:
//  * kw-EXPANDED (`big` has <= 64 channels: the high-resolution layers).  `big` is staged as three
//    column-shifted copies (kw = 0, 1, 2), 8 pixels per row each, as extra channel planes.  M runs over
//    (kw, channel) = 96 / 192 rows, so the 128-row MMA is filled, the nine taps cost three MMAs (one per
//    kh, each with its own accumulator) and nothing is staged twice.  grid.y = N tiles of `small`.
//  * TAP-GROUP (`big` has >= 128 channels).  A filter tap is a start address into the staged halo of
//    `big`; the side with fewer channels is on N; TG taps x N columns of TMEM (<= 512) per CTA.
//    grid.y = (M tile, N tile, tap group).
// Each CTA accumulates over its whole pixel range (grid.x splits the pixels) and adds its partial filter
// gradient to dw with fp32 atomics at the end.
// Warps 0-3 stage the M-side operand, 4-7 the N-side operand: global loads are issued `ds` tiles ahead as
// 16-byte cp.async copies into an fp32 staging ring, then split fp32 -> bf16 hi/lo into the operand ring.
// Warp 8 issues the MMAs; warps 0-7 drain TMEM at the end.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

struct TcwParams {
  const float* big;     // [n, bh, bw, adim]
  const float* small;   // [n, sh, sw, bdim]
  float* dw;
  int n_img, bh, bw, sh, sw, adim, bdim, stride;
  int big_is_m;         // 1: big on M, small on N; 0: small on M, big on N
  int mode_e;           // kw-expanded layout (big on M)
  int mch, nch;         // total channels on the M / N side
  int mt, nt;           // channels (rows / columns) per CTA tile on each side
  int m_tiles, n_tiles, tap_groups, tg;
  int tpr;              // tile rows (small side), tile is tpr x 8 pixels
  int tiles_h, tiles_w, ntiles;
  int stages, ds, out_ab, tmem_cols;
  float alpha;
  uint32_t m_plane, n_plane, m_bytes, stage_bytes;   // bf16 operand ring (bytes)
  uint32_t m_raw, raw_bytes;                          // fp32 staging ring: M-side bytes, slot bytes
};

constexpr int TCW_THREADS = 288;
constexpr int TCW_MAX_STAGES = 4;

// raw (fp32) halo pixels of `big` per tile, and staged pixels per plane
__host__ __device__ inline int tcw_big_raw_pixels(int tpr, int stride) { return stride == 1 ? (tpr + 2) * 10 : (2 * tpr + 1) * 17; }
__host__ __device__ inline int tcw_big_plane_pixels(int tpr, int stride, int mode_e) {
  if (mode_e) return (stride == 1 ? tpr + 2 : 2 * tpr + 1) * 8;
  return tcw_big_raw_pixels(tpr, stride);
}

__global__ void __launch_bounds__(TCW_THREADS, 1) conv_tcw_kernel(const TcwParams p) {
  extern __shared__ __align__(128) unsigned char tcw_smem[];
  __shared__ uint64_t full_m[TCW_MAX_STAGES], full_n[TCW_MAX_STAGES], empty[TCW_MAX_STAGES], done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // job decode
  int job = blockIdx.y;
  const int tgi = job % p.tap_groups;
  job /= p.tap_groups;
  const int nti = job % p.n_tiles;
  const int mti = job / p.n_tiles;
  const int tap0 = tgi * p.tg;
  const int ntap = (9 - tap0 < p.tg) ? 9 - tap0 : p.tg;
  const int m_ch0 = p.mode_e ? 0 : mti * p.mt, n_ch0 = nti * p.nt;
  const int mt_valid = (p.mch - m_ch0 < p.mt) ? p.mch - m_ch0 : p.mt;
  const int nt_valid = (p.nch - n_ch0 < p.nt) ? p.nch - n_ch0 : p.nt;
  const int qm = p.mt / 8, qn = p.nt / 8;
  unsigned char* raw_smem = tcw_smem + (size_t)p.stages * p.stage_bytes;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_m[s], 128); tc::mbar_init(&full_n[s], 128); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&done, 1);
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ============================== operand staging ====================================================
    const bool m_side = warp < 4;
    const bool is_big = m_side ? (p.big_is_m != 0) : (p.big_is_m == 0);
    const bool expand = p.mode_e && is_big;
    const int ct = tid & 127;
    const int ch_cnt = expand ? p.adim : (m_side ? p.mt : p.nt);     // channels this group stages per pixel
    const int q_cnt = ch_cnt / 8, cpp = ch_cnt / 4;                   // 16-byte bf16 vectors / fp32 chunks per pixel
    const int ch0 = expand ? 0 : (m_side ? m_ch0 : n_ch0);
    const uint32_t plane = m_side ? p.m_plane : p.n_plane;
    const float* src_base = is_big ? p.big : p.small;
    const int cdim = is_big ? p.adim : p.bdim;
    const int ih_max = is_big ? p.bh : p.sh, iw_max = is_big ? p.bw : p.sw;
    const int npx = is_big ? tcw_big_raw_pixels(p.tpr, p.stride) : p.tpr * 8;
    const uint32_t raw_off = m_side ? 0u : p.m_raw;
    const uint32_t raw_row = (uint32_t)cpp * 16u;                     // bytes of one raw pixel
    uint64_t* full = m_side ? full_m : full_n;

    // pixel ps of the raw tile -> image coordinates
    auto coords = [&](int tile, int ps, int& n, int& iy, int& ix) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      n = t / p.tiles_h;
      const int oy0 = th_ * p.tpr, ox0 = tw_ * 8;
      if (!is_big) { iy = oy0 + ps / 8; ix = ox0 + ps % 8; }
      else if (p.stride == 1) { iy = oy0 - 1 + ps / 10; ix = ox0 - 1 + ps % 10; }
      else if (expand) { iy = 2 * oy0 + ps / 17; ix = 2 * ox0 + ps % 17; }
      else { int hr = ps / 17, rem = ps % 17, par = rem >= 9; iy = 2 * oy0 + hr; ix = 2 * ox0 + 2 * (rem - 9 * par) + par; }
    };
    auto issue = [&](int tile, int slot) {
      unsigned char* rg = raw_smem + (size_t)slot * p.raw_bytes + raw_off;
      for (int ps = ct; ps < npx; ps += 128) {
        int n, iy, ix;
        coords(tile, ps, n, iy, ix);
        const bool ok = iy >= 0 && iy < ih_max && ix >= 0 && ix < iw_max;
        const float* src = ok ? src_base + (((size_t)n * ih_max + iy) * iw_max + ix) * cdim + ch0 : src_base;
        unsigned char* row = rg + (size_t)ps * raw_row;
        for (int j = 0; j < cpp; ++j) tc::cp_async16(row + (size_t)((j ^ (ps & 7)) * 16), src + j * 4, ok ? 16u : 0u);
      }
    };
    auto convert = [&](int slot, unsigned char* st) {
      const unsigned char* rg = raw_smem + (size_t)slot * p.raw_bytes + raw_off;
      for (int ps = ct; ps < npx; ps += 128) {
        const unsigned char* row = rg + (size_t)ps * raw_row;
        // destination pixel positions (up to three shifted copies in the kw-expanded layout)
        int pos[3] = {ps, -1, -1};
        if (expand) {
          if (p.stride == 1) {
            const int hr = ps / 10, hc = ps % 10;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) { const int c = hc - kw; pos[kw] = (c >= 0 && c < 8) ? hr * 8 + c : -1; }
          } else {
            const int hr = ps / 17, ic = ps % 17;
            pos[0] = (!(ic & 1) && ic <= 14) ? hr * 8 + (ic >> 1) : -1;
            pos[1] = (ic & 1) ? hr * 8 + (ic >> 1) : -1;
            pos[2] = (!(ic & 1) && ic >= 2) ? hr * 8 + (ic >> 1) - 1 : -1;
          }
        }
        for (int q = 0; q < q_cnt; ++q) {
          const float4 v0 = *reinterpret_cast<const float4*>(row + (size_t)(((2 * q) ^ (ps & 7)) * 16));
          const float4 v1 = *reinterpret_cast<const float4*>(row + (size_t)(((2 * q + 1) ^ (ps & 7)) * 16));
          uint4 h4, l4;
          tc::split2_bf16(v0.x, v0.y, h4.x, l4.x);
          tc::split2_bf16(v0.z, v0.w, h4.y, l4.y);
          tc::split2_bf16(v1.x, v1.y, h4.z, l4.z);
          tc::split2_bf16(v1.z, v1.w, h4.w, l4.w);
          if (!expand) {
            *reinterpret_cast<uint4*>(st + (size_t)q * plane + (size_t)ps * 16) = h4;
            *reinterpret_cast<uint4*>(st + (size_t)(q_cnt + q) * plane + (size_t)ps * 16) = l4;
          } else {
            // planes [split][kw][q]
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              if (pos[kw] >= 0) {
                *reinterpret_cast<uint4*>(st + (size_t)(kw * q_cnt + q) * plane + (size_t)pos[kw] * 16) = h4;
                *reinterpret_cast<uint4*>(st + (size_t)((3 + kw) * q_cnt + q) * plane + (size_t)pos[kw] * 16) = l4;
              }
          }
        }
      }
    };

    int stage = 0;
    uint32_t phase = 0;
    int pt = blockIdx.x, slot_pf = 0;
    for (int i = 0; i < p.ds; ++i) {
      if (pt < p.ntiles) { issue(pt, slot_pf); pt += gridDim.x; }
      tc::cp_async_commit();
      if (++slot_pf == p.ds) slot_pf = 0;
    }
    int slot_cv = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      if (p.ds == 1) tc::cp_async_wait<0>();
      else if (p.ds == 2) tc::cp_async_wait<1>();
      else tc::cp_async_wait<2>();
      tc::mbar_wait(&empty[stage], phase ^ 1u);
      convert(slot_cv, tcw_smem + (size_t)stage * p.stage_bytes + (m_side ? 0u : p.m_bytes));
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      if (pt < p.ntiles) { issue(pt, slot_cv); pt += gridDim.x; }
      tc::cp_async_commit();
      if (++slot_cv == p.ds) slot_cv = 0;
    }
    tc::cp_async_wait<0>();

    // ============================== drain: TMEM -> atomics into dw =====================================
    tc::mbar_wait(&done, 0);
    tc::tc_fence_after();
    const int quarter = warp & 3, half = warp >> 2;
    const int m = quarter * 32 + lane;
    const int ncols = p.mode_e ? 3 * p.m_tiles * p.nt : ntap * p.nt;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int qa = p.adim / 8;
    for (int c0 = half * 32; c0 < ncols; c0 += 64) {
      float v[32];
      tc::tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);
      const int accn = c0 / p.nt, nn0 = c0 % p.nt;
      int tap, a_or_m;
      bool row_ok;
      if (p.mode_e) {
        // accumulator (kh, t) at ((kh * m_tiles + t) * nt); row m of M tile t = group 16 t + m/8 of (kw, channel)
        const int kh = accn / p.m_tiles, t = accn % p.m_tiles;
        const int gi = 16 * t + (m >> 3);
        row_ok = gi < 3 * qa;
        tap = kh * 3 + gi / qa;
        a_or_m = (gi % qa) * 8 + (m & 7);
      } else {
        row_ok = m < mt_valid;
        tap = tap0 + accn;
        a_or_m = m_ch0 + m;
      }
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int nn = nn0 + j;
          if (nn < nt_valid) {
            const int a = p.big_is_m ? a_or_m : (n_ch0 + nn);
            const int b = p.big_is_m ? (n_ch0 + nn) : a_or_m;
            const size_t o = p.out_ab ? ((size_t)tap * p.adim + a) * p.bdim + b : ((size_t)tap * p.bdim + b) * p.adim + a;
            atomicAdd(p.dw + o, v[j] * p.alpha);
          }
        }
      }
    }
  } else {
    // ============================== MMA issue ==========================================================
    // Descriptors = per-stage base + tap offset + k-step stride, all in 16-byte units in the low word.
    // The whole warp runs the control flow (uniform registers); one elected lane issues.
    const uint32_t idesc = tc::idesc_bf16_f32(p.nt, 1, 1);
    const uint32_t stage16 = p.stage_bytes >> 4;
    const int ksteps = p.tpr / 2;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum_first = 0;
    if (p.mode_e) {
      // big on M: tile row r, tap kh -> staged row (r*S + kh) of 8 pixels
      const uint32_t big_lbo = (p.stride == 1) ? 128u : 256u;
      const uint32_t big_step16 = (p.stride == 1) ? 16u : 32u;        // two tile rows
      const uint64_t m_desc0 = tc::smem_desc(tc::smem_u32(tcw_smem), big_lbo, p.m_plane);
      const uint64_t n_desc0 = tc::smem_desc(tc::smem_u32(tcw_smem) + p.m_bytes, 128u, p.n_plane);
      const uint32_t split_lo16 = ((uint32_t)(3 * (p.adim / 8)) * p.m_plane) >> 4;
      const uint32_t n_lo16 = ((uint32_t)qn * p.n_plane) >> 4;
      const uint32_t mtile16 = (16u * p.m_plane) >> 4;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        tc::mbar_wait(&full_m[stage], phase);
        tc::mbar_wait(&full_n[stage], phase);
        tc::tc_fence_after();
        const uint64_t m_base = m_desc0 + (uint64_t)((uint32_t)stage * stage16);
        const uint64_t n_base = n_desc0 + (uint64_t)((uint32_t)stage * stage16);
        if (tc::elect_one()) {
          for (int kh = 0; kh < 3; ++kh) {
            for (int t = 0; t < p.m_tiles; ++t) {
              uint64_t a_hi = m_base + (uint64_t)((uint32_t)kh * 8u + (uint32_t)t * mtile16);
              uint64_t b_hi = n_base;
              const uint32_t d = tmem_base + (uint32_t)((kh * p.m_tiles + t) * p.nt);
              uint32_t accum = accum_first;
#pragma unroll 2
              for (int j = 0; j < ksteps; ++j) {
                tc::mma_bf16(d, a_hi, b_hi, idesc, accum);
                tc::mma_bf16(d, a_hi, b_hi + n_lo16, idesc, 1u);
                tc::mma_bf16(d, a_hi + split_lo16, b_hi, idesc, 1u);
                accum = 1u;
                a_hi += big_step16;
                b_hi += 16u;
              }
            }
          }
          tc::mma_commit(&empty[stage]);
        }
        __syncwarp();
        accum_first = 1u;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    } else {
      const uint32_t big_lbo = (p.stride == 1) ? 160u : 544u;            // next tile row of `big`
      const uint32_t big_step16 = (p.stride == 1) ? 20u : 68u;            // two tile rows of `big`
      const uint32_t m_step16 = p.big_is_m ? big_step16 : 16u, n_step16 = p.big_is_m ? 16u : big_step16;
      const uint64_t m_desc0 = tc::smem_desc(tc::smem_u32(tcw_smem), p.big_is_m ? big_lbo : 128u, p.m_plane);
      const uint64_t n_desc0 = tc::smem_desc(tc::smem_u32(tcw_smem) + p.m_bytes, p.big_is_m ? 128u : big_lbo, p.n_plane);
      const uint32_t m_lo16 = ((uint32_t)qm * p.m_plane) >> 4, n_lo16 = ((uint32_t)qn * p.n_plane) >> 4;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        tc::mbar_wait(&full_m[stage], phase);
        tc::mbar_wait(&full_n[stage], phase);
        tc::tc_fence_after();
        const uint64_t m_base = m_desc0 + (uint64_t)((uint32_t)stage * stage16);
        const uint64_t n_base = n_desc0 + (uint64_t)((uint32_t)stage * stage16);
        if (tc::elect_one()) {
          for (int tl = 0; tl < ntap; ++tl) {
            const int tap = tap0 + tl, kh = tap / 3, kw = tap % 3;
            const uint32_t tap16 = (p.stride == 1) ? (uint32_t)(kh * 10 + kw) : (uint32_t)(kh * 17 + (kw & 1) * 9 + (kw >> 1));
            uint64_t a_hi = m_base + (uint64_t)(p.big_is_m ? tap16 : 0u);
            uint64_t b_hi = n_base + (uint64_t)(p.big_is_m ? 0u : tap16);
            const uint32_t d = tmem_base + (uint32_t)(tl * p.nt);
            uint32_t accum = accum_first;
#pragma unroll 2
            for (int j = 0; j < ksteps; ++j) {
              tc::mma_bf16(d, a_hi, b_hi, idesc, accum);
              tc::mma_bf16(d, a_hi, b_hi + n_lo16, idesc, 1u);
              tc::mma_bf16(d, a_hi + m_lo16, b_hi, idesc, 1u);
              accum = 1u;
              a_hi += m_step16;
              b_hi += n_step16;
            }
          }
          tc::mma_commit(&empty[stage]);
        }
        __syncwarp();
        accum_first = 1u;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    if (tc::elect_one()) tc::mma_commit(&done);
    __syncwarp();
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
