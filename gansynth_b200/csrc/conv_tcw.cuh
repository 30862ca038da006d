// tcgen05 / TMEM filter-gradient ("W form") of the 3x3 convolutions, bf16x3, for sm_100a.
//
//   dw(tap, a, b) = alpha * sum_{n,oy,ox} big[n, oy*S+kh-pb, ox*S+kw-pb, a] * small[n, oy, ox, b]
//
// Per tap this is a GEMM whose contraction index is the PIXEL.  Both operands sit in shared memory in the
// same layout the forward kernels use (16-byte channel vectors, pixel-major planes), which for this GEMM
// is the MN-major SWIZZLE_NONE core-matrix layout: channel chunks along M / N (stride = plane), pixels
// along K (8 consecutive pixels = one image-row segment, next segment = next tile row).  A filter tap is
// a different start address into the staged halo of `big`, so one staged tile feeds all taps.
// The side with fewer channels is put on N (<= 128 per CTA); TG taps x N columns of TMEM (<= 512) are
// accumulated over the CTA's whole pixel range and added to dw with fp32 atomics at the end.
// grid = (pixel splits, jobs), job = (M tile, N tile, tap group).
//
// Thin layers (<= 64 channels on the `big` side) use the kw-EXPANDED mode instead: `big` is staged three
// times, shifted by the column offset of kw = 0, 1, 2, as extra channel planes.  M then runs over
// (kw, channel) = 96 or 192 rows, so a 128-row MMA is filled, the nine taps cost three MMAs (one per kh,
// each with its own accumulator) and nothing is re-staged per tap group.
// Warps 0-3 stage the M-side operand, 4-7 the N-side operand (fp32 -> bf16 hi/lo), warp 8 issues MMAs;
// warps 0-7 drain TMEM at the end.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

struct TcwParams {
  const float* big;     // [n, bh, bw, adim]
  const float* small;   // [n, sh, sw, bdim]
  float* dw;
  int n_img, bh, bw, sh, sw, adim, bdim, stride;
  int big_is_m;         // 1: big on M, small on N; 0: small on M, big on N
  int mode_e;           // kw-expanded mode (big on M, M = 3 * adim rows in m_tiles tiles of 128)
  int mch, nch;         // total channels on the M / N side
  int mt, nt;           // channels per CTA tile on each side (mt <= 128, nt <= 128, multiples of 8)
  int m_tiles, n_tiles, tap_groups, tg;
  int tpr;              // tile rows (small side), tile is tpr x 8 pixels
  int tiles_h, tiles_w, ntiles;
  int stages, out_ab, tmem_cols;
  float alpha;
  uint32_t m_plane, n_plane, m_bytes, stage_bytes;   // bytes
};

constexpr int TCW_THREADS = 288;
constexpr int TCW_MAX_STAGES = 4;

// pixel count of the staged tile of `big` / `small`
__host__ __device__ inline int tcw_big_pixels(int tpr, int stride) { return stride == 1 ? (tpr + 2) * 10 : (2 * tpr + 1) * 17; }

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
  const int m_ch0 = mti * p.mt, n_ch0 = nti * p.nt;
  const int mt_valid = (p.mch - m_ch0 < p.mt) ? p.mch - m_ch0 : p.mt;
  const int nt_valid = (p.nch - n_ch0 < p.nt) ? p.nch - n_ch0 : p.nt;
  const int qm = p.mt / 8, qn = p.nt / 8;

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
    const int ct = tid & 127;
    const bool expand = p.mode_e && is_big;          // kw-expanded staging of `big`
    const int q_cnt = expand ? p.adim / 8 : (m_side ? qm : qn);
    const int ch0 = expand ? 0 : (m_side ? m_ch0 : n_ch0);
    const int ch_valid = expand ? p.adim : (m_side ? mt_valid : nt_valid);
    const int shift1 = (p.stride == 1) ? 1 : 9, shift2 = (p.stride == 1) ? 2 : 1;   // staged-pixel shift of kw = 1, 2
    const uint32_t plane = m_side ? p.m_plane : p.n_plane;
    const float* src_base = is_big ? p.big : p.small;
    const int cdim = is_big ? p.adim : p.bdim;
    const int ih_max = is_big ? p.bh : p.sh, iw_max = is_big ? p.bw : p.sw;
    const int npx = is_big ? tcw_big_pixels(p.tpr, p.stride) : p.tpr * 8;
    uint64_t* full = m_side ? full_m : full_n;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw_ = t % p.tiles_w;
      t /= p.tiles_w;
      const int th_ = t % p.tiles_h;
      const int n = t / p.tiles_h;
      const int oy0 = th_ * p.tpr, ox0 = tw_ * 8;
      tc::mbar_wait(&empty[stage], phase ^ 1u);
      unsigned char* st = tcw_smem + (size_t)stage * p.stage_bytes + (m_side ? 0u : p.m_bytes);
      for (int ps = ct; ps < npx; ps += 128) {
        int iy, ix;
        if (!is_big) { iy = oy0 + ps / 8; ix = ox0 + ps % 8; }
        else if (p.stride == 1) { iy = oy0 - 1 + ps / 10; ix = ox0 - 1 + ps % 10; }
        else { int hr = ps / 17, rem = ps % 17, par = rem >= 9; iy = 2 * oy0 + hr; ix = 2 * ox0 + 2 * (rem - 9 * par) + par; }
        const bool ok = iy >= 0 && iy < ih_max && ix >= 0 && ix < iw_max;
        const float* src = src_base + (((size_t)n * ih_max + iy) * iw_max + ix) * cdim + ch0;
        for (int q = 0; q < q_cnt; ++q) {
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (ok && q * 8 < ch_valid) {
            v0 = __ldg(reinterpret_cast<const float4*>(src + q * 8));
            v1 = __ldg(reinterpret_cast<const float4*>(src + q * 8 + 4));
          }
          const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          __nv_bfloat16 hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) tc::split_bf16(f[e], hi[e], lo[e]);
          uint4 h4, l4;
          h4.x = tc::pack_bf16(hi[0], hi[1]); h4.y = tc::pack_bf16(hi[2], hi[3]);
          h4.z = tc::pack_bf16(hi[4], hi[5]); h4.w = tc::pack_bf16(hi[6], hi[7]);
          l4.x = tc::pack_bf16(lo[0], lo[1]); l4.y = tc::pack_bf16(lo[2], lo[3]);
          l4.z = tc::pack_bf16(lo[4], lo[5]); l4.w = tc::pack_bf16(lo[6], lo[7]);
          if (!expand) {
            *reinterpret_cast<uint4*>(st + (size_t)q * plane + (size_t)ps * 16) = h4;
            *reinterpret_cast<uint4*>(st + (size_t)(q_cnt + q) * plane + (size_t)ps * 16) = l4;
          } else {
            // planes [split][kw][q]; plane kw holds the tile shifted left by the column offset of tap kw
            const size_t lo0 = (size_t)3 * q_cnt;
            *reinterpret_cast<uint4*>(st + (size_t)q * plane + (size_t)ps * 16) = h4;
            *reinterpret_cast<uint4*>(st + (lo0 + q) * plane + (size_t)ps * 16) = l4;
            if (ps >= shift1) {
              *reinterpret_cast<uint4*>(st + (size_t)(q_cnt + q) * plane + (size_t)(ps - shift1) * 16) = h4;
              *reinterpret_cast<uint4*>(st + (lo0 + q_cnt + q) * plane + (size_t)(ps - shift1) * 16) = l4;
            }
            if (ps >= shift2) {
              *reinterpret_cast<uint4*>(st + (size_t)(2 * q_cnt + q) * plane + (size_t)(ps - shift2) * 16) = h4;
              *reinterpret_cast<uint4*>(st + (lo0 + 2 * q_cnt + q) * plane + (size_t)(ps - shift2) * 16) = l4;
            }
          }
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
    // ============================== drain: TMEM -> atomics into dw =====================================
    tc::mbar_wait(&done, 0);
    tc::tc_fence_after();
    const int quarter = warp & 3, half = warp >> 2;
    const int m = quarter * 32 + lane;
    const int ncols = p.mode_e ? 3 * p.m_tiles * p.nt : ntap * p.nt;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    if (p.mode_e) {
      // columns: accumulator (kh, t) at ((kh * m_tiles + t) * nt); row m of M tile t = group 16 t + m/8 of (kw, channel)
      const int qa = p.adim / 8;
      for (int c0 = half * 32; c0 < ncols; c0 += 64) {
        float v[32];
        tc::tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);
        const int accn = c0 / p.nt, nn0 = c0 % p.nt;
        const int kh = accn / p.m_tiles, t = accn % p.m_tiles;
        const int gi = 16 * t + (m >> 3);
        if (gi < 3 * qa) {
          const int kw = gi / qa, a = (gi % qa) * 8 + (m & 7);
          const int tap = kh * 3 + kw;
#pragma unroll 4
          for (int j = 0; j < 32; ++j) {
            const int nn = nn0 + j;
            if (nn < nt_valid) {
              const int b = n_ch0 + nn;
              const size_t o = p.out_ab ? ((size_t)tap * p.adim + a) * p.bdim + b : ((size_t)tap * p.bdim + b) * p.adim + a;
              atomicAdd(p.dw + o, v[j] * p.alpha);
            }
          }
        }
      }
    } else
    for (int c0 = half * 32; c0 < ncols; c0 += 64) {
      float v[32];
      tc::tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);
      if (m < mt_valid) {
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const int col = c0 + j;
          const int tl = col / p.nt, nn = col % p.nt;
          if (col < ncols && nn < nt_valid) {
            const int a = p.big_is_m ? (m_ch0 + m) : (n_ch0 + nn);
            const int b = p.big_is_m ? (n_ch0 + nn) : (m_ch0 + m);
            const int tap = tap0 + tl;
            const size_t o = p.out_ab ? ((size_t)tap * p.adim + a) * p.bdim + b : ((size_t)tap * p.bdim + b) * p.adim + a;
            atomicAdd(p.dw + o, v[j] * p.alpha);
          }
        }
      }
    }
  } else if (lane == 0) {
    // ============================== MMA issue ==========================================================
    // Descriptors = per-stage base + tap offset + k-step stride, all in 16-byte units in the low word.
    const uint32_t idesc = tc::idesc_bf16_f32(p.nt, 1, 1);
    const uint32_t big_lbo = (p.stride == 1) ? 160u : 544u;            // next tile row of `big`
    const uint32_t big_step16 = (p.stride == 1) ? 20u : 68u;            // two tile rows of `big`, 16-byte units
    const uint32_t m_step16 = p.big_is_m ? big_step16 : 16u, n_step16 = p.big_is_m ? 16u : big_step16;
    const uint64_t m_desc0 = tc::smem_desc(tc::smem_u32(tcw_smem), p.big_is_m ? big_lbo : 128u, p.m_plane);
    const uint64_t n_desc0 = tc::smem_desc(tc::smem_u32(tcw_smem) + p.m_bytes, p.big_is_m ? 128u : big_lbo, p.n_plane);
    const uint32_t stage16 = p.stage_bytes >> 4;
    const uint32_t m_lo16 = ((uint32_t)qm * p.m_plane) >> 4, n_lo16 = ((uint32_t)qn * p.n_plane) >> 4;
    const int ksteps = p.tpr / 2;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum_first = 0;
    if (p.mode_e) {
      const uint32_t split_lo16 = ((uint32_t)(3 * (p.adim / 8)) * p.m_plane) >> 4;   // hi planes -> lo planes of big
      const uint32_t row16 = (p.stride == 1) ? 10u : 17u;                            // one staged row of `big`
      const uint32_t mtile16 = (16u * p.m_plane) >> 4;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        tc::mbar_wait(&full_m[stage], phase);
        tc::mbar_wait(&full_n[stage], phase);
        tc::tc_fence_after();
        const uint64_t m_base = m_desc0 + (uint64_t)((uint32_t)stage * stage16);
        const uint64_t n_base = n_desc0 + (uint64_t)((uint32_t)stage * stage16);
        for (int kh = 0; kh < 3; ++kh) {
          for (int t = 0; t < p.m_tiles; ++t) {
            uint64_t a_hi = m_base + (uint64_t)((uint32_t)kh * row16 + (uint32_t)t * mtile16);
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
        accum_first = 1u;
        tc::mma_commit(&empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    } else
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      tc::mbar_wait(&full_m[stage], phase);
      tc::mbar_wait(&full_n[stage], phase);
      tc::tc_fence_after();
      const uint64_t m_base = m_desc0 + (uint64_t)((uint32_t)stage * stage16);
      const uint64_t n_base = n_desc0 + (uint64_t)((uint32_t)stage * stage16);
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
      accum_first = 1u;
      tc::mma_commit(&empty[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
    tc::mma_commit(&done);
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
