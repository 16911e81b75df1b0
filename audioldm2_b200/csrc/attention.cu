// K5: fused softmax(scale * Q K^T + mask) V, head_dim 32, flash-style (no N x N score matrix in HBM;
// the reference materialises it: attention.py:354-366).
//
// attention_tc_kernel (tcgen05): one CTA = 128 queries of one (batch, head); keys are streamed in
// tiles of 64.  Operands are the bf16 hi/lo planes written by the projection GEMMs
// (ALDM_OUT_QKV): Q, K row-major, V already transposed (keys contiguous), so every operand tile is
// a plain cp.async copy into the same 128-byte-swizzled K-major layout the GEMM uses.
//   S = Q K^T       : q_hi k_hi^T + q_hi k_lo^T + q_lo k_hi^T, 2 K-steps of 16 each      -> TMEM (6 UMMAs)
//   softmax          : 128 threads, one query row each (TMEM lane == row), online max/sum in the
//                      log2 domain, P split to bf16 hi/lo and written to shared memory as the A operand
//   O_tile = P V     : 3 passes x 4 K-steps, N = 32                                    -> TMEM (12 UMMAs)
//   O += rescaled O_tile in registers (no TMEM read-modify-write)
// Warp roles: 0-3 softmax/epilogue, 4 loader (cp.async + mbarrier), 5 TMEM alloc + MMA issue.
// ~97 KB shared memory and 128 TMEM columns per CTA -> two CTAs per SM overlap each other's MMA
// and softmax phases.
//
// attention_simt_kernel: CUDA-core checker on the same operands (validation only).
// softmax_rows_kernel: row softmax for the VAE AttnBlock (model.py:216-217).
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace aldm {

static constexpr int ATT_D = 32;

// ------------------------------------------------------------------------------------------------
// tcgen05 flash attention
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

namespace atc {
constexpr int QT = 128, KT = 64;
// NS = number of K/V stages.  The profile of the 2-stage kernel showed the softmax warps idle 43% of the time
// waiting for S: a stage is only released after P V(it), so K(it+2) was requested one tile period before it was
// needed, less than the L2 latency under load.  Three stages give two periods of prefetch distance.  They fit in
// the same 97 KB (two CTAs per SM) because Q is stored once as [q_hi | q_lo] rows: every K=16 MMA step takes its
// own descriptor, so the three product terms just use different 32-byte offsets inside the 128-byte swizzled rows
// (an earlier layout kept a second [q_hi | q_hi] copy to pair with [k_hi | k_lo]).
template <int NS>
struct Cfg {
  static constexpr int QA = 0;                           // [128][128B]  q_hi | q_lo
  static constexpr int KB = QA + QT * 128;               // NS x [64][128B]   k_hi | k_lo
  static constexpr int VT = KB + NS * KT * 128;          // NS x {hi,lo} x [32][128B]
  static constexpr int PP = VT + NS * 2 * ATT_D * 128;   // {hi,lo} x [128][128B]
  static constexpr int BAR = PP + 2 * QT * 128;
  static constexpr int SMEM = BAR + 128 + 1024;          // + barriers + round-up slack for the 1024-byte tile alignment
};
constexpr int TMEM_COLS = 256;                 // S double buffer: cols [0,64) / [64,128); O_tile: cols [128,160)
}  // namespace atc

template <int NS>
__global__ void __launch_bounds__(192, 2) attention_tc_kernel(const __grid_constant__ aldm_attn_desc d) {
  using namespace atc;
  using L = Cfg<NS>;
  constexpr int QA = L::QA, KB = L::KB, VT = L::VT, PP = L::PP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar = base + L::BAR;
  const uint32_t q_full = bar, kv_full0 = bar + 8, kv_empty0 = kv_full0 + 8 * NS, s_full0 = kv_empty0 + 8 * NS,
                 p_full = s_full0 + 16, o_full = p_full + 8, tmem_slot = o_full + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int bkv = d.kv_bmod > 0 ? b % d.kv_bmod : b;
  const int nt = (d.Nk + KT - 1) / KT;

  if (tid == 0) {
    // one arrival per WARP everywhere (per-thread arrivals on one mbarrier word serialise in the smem atomic unit)
    mbar_init(q_full, 32);
    for (int i = 0; i < NS; ++i) { mbar_init(kv_full0 + 8 * i, 32); mbar_init(kv_empty0 + 8 * i, 1); }
    mbar_init(s_full0, 1); mbar_init(s_full0 + 8, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (tmem_slot - base));
  const uint32_t tmem_S = tmem, tmem_O = tmem + 128;
  pdl_wait();

  if (warp < 4) {
    // =============================== softmax + output ===============================
    const int row = tid, q = q0 + row;
    const float sl2 = d.scale * 1.4426950408889634f;
    const uint32_t trow = (uint32_t)(warp * 32) << 16;
    const uint32_t swz = (uint32_t)(row & 7);
    float o[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o[i] = 0.f;
    float mrun = -INFINITY, lrun = 0.f;
    const float* mrow = d.mask ? d.mask + (long long)bkv * d.Nk : nullptr;
    for (int it = 0; it < nt; ++it) {
      const int k0 = it * KT;
      mbar_wait(s_full0 + 8 * (it & 1), (it >> 1) & 1);
      tc_fence_after();
      float s[KT];
      tmem_ld32(tmem_S + (it & 1) * KT + trow, reinterpret_cast<uint32_t*>(s));
      tmem_ld32(tmem_S + (it & 1) * KT + trow + 32, reinterpret_cast<uint32_t*>(s + 32));
      tmem_ld_wait();
      float mnew, corr, psum = 0.f;
      if (k0 + KT <= d.Nk && !mrow) {
        // interior tile, no mask: max on the raw scores (sl2 > 0), one FFMA + one MUFU per element
        float tmax = s[0];
#pragma unroll
        for (int j = 1; j < KT; ++j) tmax = fmaxf(tmax, s[j]);
        mnew = fmaxf(mrun, tmax * sl2);
        corr = ex2_approx(mrun - mnew);                  // mrun = -inf on the first tile -> 0
#pragma unroll
        for (int j = 0; j < KT; ++j) {
          const float p = ex2_approx(fmaf(s[j], sl2, -mnew));
          s[j] = p;
          psum += p;
        }
      } else {
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
          const int key = k0 + j;
          float v = s[j] * sl2;
          if (key >= d.Nk) v = -INFINITY;                               // beyond the key range: excluded
          else if (mrow && __ldg(mrow + key) != 1.0f) v = -FLT_MAX;     // masked_fill(-finfo.max), attention.py:356-360
          s[j] = v;
          tmax = fmaxf(tmax, v);
        }
        mnew = fmaxf(mrun, tmax);
        corr = (mrun == -INFINITY) ? 0.f : ex2_approx(mrun - mnew);
#pragma unroll
        for (int j = 0; j < KT; ++j) {
          const float p = (s[j] == -INFINITY) ? 0.f : ex2_approx(s[j] - mnew);
          s[j] = p;
          psum += p;
        }
      }
      lrun = lrun * corr + psum;
      mrun = mnew;
      // Deferred accumulation: fold in O_tile of the PREVIOUS key tile (its P V product has long finished
      // while this tile's probabilities were computed), then rescale to the new running maximum.  Waiting
      // for o_full(it-1) here also guarantees the tensor core is done reading the P buffer we overwrite next.
      if (it > 0) {
        mbar_wait(o_full, (it - 1) & 1);
        tc_fence_after();
        float ot[ATT_D];
        tmem_ld32(tmem_O + trow, reinterpret_cast<uint32_t*>(ot));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < ATT_D; ++i) o[i] = (o[i] + ot[i]) * corr;
      }
      uint8_t* ph = sm + PP + row * 128;
#pragma unroll
      for (int c = 0; c < KT / 8; ++c) {
        uint4 hi, lo;
        split8(s + c * 8, hi, lo);
        const uint32_t off = ((uint32_t)c ^ swz) << 4;
        *reinterpret_cast<uint4*>(ph + off) = hi;
        *reinterpret_cast<uint4*>(ph + QT * 128 + off) = lo;
      }
      fence_proxy_async();          // P (generic-proxy stores) -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    {
      mbar_wait(o_full, (nt - 1) & 1);
      tc_fence_after();
      float ot[ATT_D];
      tmem_ld32(tmem_O + trow, reinterpret_cast<uint32_t*>(ot));
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < ATT_D; ++i) o[i] += ot[i];
    }
    tc_fence_before();
    if (q < d.Nq) {
      const float inv = 1.0f / lrun;
#pragma unroll
      for (int i = 0; i < ATT_D; ++i) o[i] *= inv;
      const long long orow = (long long)b * d.Nq + q;
      __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(d.out_hi) + orow * d.ldo + h * ATT_D;
      __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(d.out_lo) + orow * d.ldo + h * ATT_D;
#pragma unroll
      for (int i = 0; i < ATT_D; i += 8) {
        uint4 hh, ll;
        split8(o + i, hh, ll);
        *reinterpret_cast<uint4*>(hp + i) = hh;
        *reinterpret_cast<uint4*>(lp + i) = ll;
      }
    }
  } else if (warp == 4) {
    // =============================== loader ===============================
    const __nv_bfloat16* qh = reinterpret_cast<const __nv_bfloat16*>(d.q_hi);
    const __nv_bfloat16* ql = reinterpret_cast<const __nv_bfloat16*>(d.q_lo);
    const __nv_bfloat16* kh = reinterpret_cast<const __nv_bfloat16*>(d.k_hi);
    const __nv_bfloat16* kl = reinterpret_cast<const __nv_bfloat16*>(d.k_lo);
    const __nv_bfloat16* vh = reinterpret_cast<const __nv_bfloat16*>(d.vt_hi);
    const __nv_bfloat16* vl = reinterpret_cast<const __nv_bfloat16*>(d.vt_lo);
    // Q: row r chunk c <- (c < 4 ? q_hi : q_lo) chunk (c & 3)
    for (int idx = lane; idx < QT * 8; idx += 32) {
      const int r = idx >> 3, c = idx & 7;
      const bool ok = q0 + r < d.Nq;
      const long long off = ok ? ((long long)b * d.Nq + q0 + r) * d.ldq + d.q_col + h * ATT_D + (c & 3) * 8 : 0;
      cp_async_16(base + QA + r * 128 + ((uint32_t)(c ^ (r & 7)) << 4), (c < 4 ? qh : ql) + off, ok ? 16u : 0u);
    }
    cp_async_mbar_arrive_noinc(q_full);
    // K/V tiles: every lane owns one 16-byte chunk column c and rows r0 + 4i; all row bases are hoisted out of
    // the tile loop (the loader is a single warp: per-element 64-bit index arithmetic was the bottleneck)
    const int c = lane & 7, r0 = lane >> 3;
    const __nv_bfloat16* kcol = (c < 4 ? kh : kl) + (long long)bkv * d.Nk * d.ldk + d.k_col + h * ATT_D + (c & 3) * 8;
    const long long vrow0 = ((long long)(bkv * d.heads + h) * ATT_D + r0) * d.ld_t + c * 8;
    const long long kstep = 4ll * d.ldk, vstep = 4ll * d.ld_t;
    for (int it = 0, s = 0, ph = 1; it < nt; ++it) {
      const int k0 = it * KT;
      mbar_wait(kv_empty0 + 8 * s, ph);
      const uint32_t kb = base + KB + s * (KT * 128);
      const __nv_bfloat16* kp = kcol + (long long)(k0 + r0) * d.ldk;
#pragma unroll
      for (int i = 0; i < KT / 4; ++i) {
        const int r = r0 + 4 * i;
        const bool ok = k0 + r < d.Nk;
        cp_async_16(kb + r * 128 + ((uint32_t)(c ^ (r & 7)) << 4), ok ? kp + i * kstep : kcol, ok ? 16u : 0u);
      }
      const uint32_t vb = base + VT + s * (2 * ATT_D * 128);
      const bool vok = k0 + c * 8 < d.Nk;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int pl = i >> 3, r = r0 + 4 * (i & 7);
        const __nv_bfloat16* vp = (pl ? vl : vh) + vrow0 + (i & 7) * vstep + k0;
        cp_async_16(vb + pl * (ATT_D * 128) + r * 128 + ((uint32_t)(c ^ (r & 7)) << 4), vok ? vp : vh, vok ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(kv_full0 + 8 * s);
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  } else {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc_bf16(128, KT), idO = umma_idesc_bf16(128, ATT_D);
      const uint64_t dQ = umma_desc_sw128(base + QA);
      const uint64_t dPh = umma_desc_sw128(base + PP), dPl = umma_desc_sw128(base + PP + QT * 128);
      auto issue_S = [&](int t) {      // S(t) = Q K(t)^T into TMEM buffer t & 1
        const int st = t % NS;
        mbar_wait(kv_full0 + 8 * st, (t / NS) & 1);      // returns at once when the caller has already seen it complete
        tc_fence_after();
        const uint64_t dK = umma_desc_sw128(base + KB + st * (KT * 128));
        const uint32_t tS = tmem_S + (t & 1) * KT;
        // descriptor address units are 16 bytes: +2 = next K-step of 16 bf16, +4 = the lo half of the row
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) umma_bf16(tS, dQ + 4 + 2 * ks, dK + 2 * ks, idS, ks > 0);     // q_lo k_hi
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) umma_bf16(tS, dQ + 2 * ks, dK + 4 + 2 * ks, idS, 1);          // q_hi k_lo
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) umma_bf16(tS, dQ + 2 * ks, dK + 2 * ks, idS, 1);              // q_hi k_hi
        umma_commit(s_full0 + 8 * (t & 1));
      };
      mbar_wait(q_full, 0);
      issue_S(0);
      for (int it = 0; it < nt; ++it) {
        const int s = it % NS;
        // S(it+1) runs on the tensor core while the softmax warps work on S(it) (its TMEM buffer was drained
        // before p_full(it-1), which this thread has already observed) -- but only if K(it+1) has landed:
        // otherwise P V(it) goes first so that its stage is released and the loader keeps prefetching.
        bool s_next = it + 1 >= nt;
        if (!s_next && mbar_test_wait(kv_full0 + 8 * ((it + 1) % NS), ((it + 1) / NS) & 1)) { issue_S(it + 1); s_next = true; }
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        const uint64_t dVh = umma_desc_sw128(base + VT + s * (2 * ATT_D * 128));
        const uint64_t dVl = umma_desc_sw128(base + VT + s * (2 * ATT_D * 128) + ATT_D * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_bf16(tmem_O, dPl + 2 * ks, dVh + 2 * ks, idO, ks > 0);
          umma_bf16(tmem_O, dPh + 2 * ks, dVl + 2 * ks, idO, 1);
          umma_bf16(tmem_O, dPh + 2 * ks, dVh + 2 * ks, idO, 1);
        }
        umma_commit(o_full);
        umma_commit(kv_empty0 + 8 * s);
        if (!s_next) issue_S(it + 1);
      }
      pdl_launch();     // last P V issued: schedule the next kernel's blocks under this CTA's tail
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------
// Short key sets (cross-attention to the 8-token CLAP/GPT-2 and 32-token T5 contexts): Nk <= 32.
// The tensor-core kernel pays its whole fixed cost (TMEM allocation, three mbarrier hand-overs, a 64-key
// tile that is mostly padding) for ~0.1 GFLOP: 37 us per launch in the step's launch list, 32 launches
// per DDIM step.  Here one thread owns one query, K and V of the (batch, head) sit in shared memory as
// fp32 (hi + lo is exact in fp32), and the 2 x Nk x 32 FMAs per query run on the CUDA cores.
// ------------------------------------------------------------------------------------------------
template <int NKT>
__global__ void __launch_bounds__(128) attention_short_kernel(const __grid_constant__ aldm_attn_desc d) {
  __shared__ __align__(16) float sk[NKT][ATT_D];
  __shared__ __align__(16) float sv[NKT][ATT_D];
  __shared__ int sstate[NKT];        // 0 = attend, 1 = masked (-FLT_MAX fill), 2 = beyond Nk
  const int tid = threadIdx.x;
  const int b = blockIdx.z, h = blockIdx.y;
  const int bkv = d.kv_bmod > 0 ? b % d.kv_bmod : b;
  pdl_wait();
  {
    const __nv_bfloat16* kh = reinterpret_cast<const __nv_bfloat16*>(d.k_hi);
    const __nv_bfloat16* kl = reinterpret_cast<const __nv_bfloat16*>(d.k_lo);
    const __nv_bfloat16* vh = reinterpret_cast<const __nv_bfloat16*>(d.vt_hi);
    const __nv_bfloat16* vl = reinterpret_cast<const __nv_bfloat16*>(d.vt_lo);
    for (int idx = tid; idx < NKT * ATT_D; idx += 128) {
      const int key = idx / ATT_D, dim = idx % ATT_D;
      float kv = 0.f, vv = 0.f;
      if (key < d.Nk) {
        const long long ki = ((long long)bkv * d.Nk + key) * d.ldk + d.k_col + h * ATT_D + dim;
        const long long vi = ((long long)(bkv * d.heads + h) * ATT_D + dim) * d.ld_t + key;
        kv = __bfloat162float(kh[ki]) + __bfloat162float(kl[ki]);
        vv = __bfloat162float(vh[vi]) + __bfloat162float(vl[vi]);
      }
      sk[key][dim] = kv;
      sv[key][dim] = vv;
    }
    if (tid < NKT) sstate[tid] = tid >= d.Nk ? 2 : ((d.mask && __ldg(d.mask + (long long)bkv * d.Nk + tid) != 1.0f) ? 1 : 0);
  }
  __syncthreads();
  pdl_launch();
  const int q = blockIdx.x * 128 + tid;
  if (q >= d.Nq) return;
  float qv[ATT_D];
  {
    const long long qi = ((long long)b * d.Nq + q) * d.ldq + d.q_col + h * ATT_D;
    const uint4* ph = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.q_hi) + qi);
    const uint4* pl = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.q_lo) + qi);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 a = __ldg(ph + c), l = __ldg(pl + c);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {      // bf16 -> fp32 is a 16-bit shift
        qv[c * 8 + 2 * e] = __uint_as_float(aw[e] << 16) + __uint_as_float(lw[e] << 16);
        qv[c * 8 + 2 * e + 1] = __uint_as_float(aw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
      }
    }
  }
  const float sl2 = d.scale * 1.4426950408889634f;
  float sc[NKT];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < NKT; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int dd = 0; dd < ATT_D; dd += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(&sk[k][dd]);
      acc = fmaf(qv[dd], kk.x, acc); acc = fmaf(qv[dd + 1], kk.y, acc);
      acc = fmaf(qv[dd + 2], kk.z, acc); acc = fmaf(qv[dd + 3], kk.w, acc);
    }
    const int st = sstate[k];
    const float v = st == 0 ? acc * sl2 : (st == 1 ? -FLT_MAX : -INFINITY);   // masked_fill(-finfo.max), attention.py:356-360
    sc[k] = v;
    mx = fmaxf(mx, v);
  }
  float l = 0.f;
#pragma unroll
  for (int k = 0; k < NKT; ++k) {
    const float pk = sc[k] == -INFINITY ? 0.f : ex2_approx(sc[k] - mx);
    sc[k] = pk;
    l += pk;
  }
  float o[ATT_D];
#pragma unroll
  for (int i = 0; i < ATT_D; ++i) o[i] = 0.f;
#pragma unroll
  for (int k = 0; k < NKT; ++k) {
#pragma unroll
    for (int dd = 0; dd < ATT_D; dd += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(&sv[k][dd]);
      o[dd] = fmaf(sc[k], vv.x, o[dd]); o[dd + 1] = fmaf(sc[k], vv.y, o[dd + 1]);
      o[dd + 2] = fmaf(sc[k], vv.z, o[dd + 2]); o[dd + 3] = fmaf(sc[k], vv.w, o[dd + 3]);
    }
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int i = 0; i < ATT_D; ++i) o[i] *= inv;
  const long long orow = (long long)b * d.Nq + q;
  __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(d.out_hi) + orow * d.ldo + h * ATT_D;
  __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(d.out_lo) + orow * d.ldo + h * ATT_D;
#pragma unroll
  for (int i = 0; i < ATT_D; i += 8) {
    uint4 hh, ll;
    split8(o + i, hh, ll);
    *reinterpret_cast<uint4*>(hp + i) = hh;
    *reinterpret_cast<uint4*>(lp + i) = ll;
  }
}

// ------------------------------------------------------------------------------------------------
// CUDA-core checker on the same plane operands: one thread per query, fp32
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_simt_kernel(const __grid_constant__ aldm_attn_desc d) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= d.Nq) return;
  const int bkv = d.kv_bmod > 0 ? b % d.kv_bmod : b;
  auto ld = [](const void* hi, const void* lo, long long i) {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(hi)[i]) +
           __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(lo)[i]);
  };
  float q[ATT_D], o[ATT_D];
  for (int i = 0; i < ATT_D; ++i) {
    q[i] = ld(d.q_hi, d.q_lo, ((long long)b * d.Nq + qi) * d.ldq + d.q_col + h * ATT_D + i) * d.scale;
    o[i] = 0.f;
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int k = 0; k < d.Nk; ++k) {
    float s = 0.f;
    const long long kr = ((long long)bkv * d.Nk + k) * d.ldk + d.k_col + h * ATT_D;
    for (int i = 0; i < ATT_D; ++i) s = fmaf(q[i], ld(d.k_hi, d.k_lo, kr + i), s);
    if (d.mask && d.mask[(long long)bkv * d.Nk + k] != 1.0f) s = -FLT_MAX;
    const float mnew = fmaxf(mrun, s);
    const float corr = (mrun == -INFINITY) ? 0.f : expf(mrun - mnew);
    const float p = expf(s - mnew);
    lrun = lrun * corr + p;
    for (int i = 0; i < ATT_D; ++i)
      o[i] = o[i] * corr + p * ld(d.vt_hi, d.vt_lo, ((long long)(bkv * d.heads + h) * ATT_D + i) * d.ld_t + k);
    mrun = mnew;
  }
  const long long orow = (long long)b * d.Nq + qi;
  __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(d.out_hi) + orow * d.ldo + h * ATT_D;
  __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(d.out_lo) + orow * d.ldo + h * ATT_D;
  for (int i = 0; i < ATT_D; ++i) {
    const float v = o[i] / lrun;
    const __nv_bfloat16 hh = __float2bfloat16_rn(v);
    hp[i] = hh;
    lp[i] = __float2bfloat16_rn(v - __bfloat162float(hh));
  }
}

int attention_launch(const aldm_attn_desc& d, cudaStream_t st) {
  ALDM_REQUIRE(d.q_hi && d.q_lo && d.k_hi && d.k_lo && d.vt_hi && d.vt_lo && d.out_hi && d.out_lo, ALDM_E_ARG,
               "attention: null pointer");
  ALDM_REQUIRE(d.B > 0 && d.heads > 0 && d.Nq > 0 && d.Nk > 0, ALDM_E_SHAPE, "attention: B=%d heads=%d Nq=%d Nk=%d", d.B,
               d.heads, d.Nq, d.Nk);
  ALDM_REQUIRE(d.ldq % 8 == 0 && d.ldk % 8 == 0 && d.ld_t % 8 == 0 && d.ldo % 8 == 0 && d.q_col % 8 == 0 && d.k_col % 8 == 0,
               ALDM_E_ALIGN, "attention: leading dims / column offsets must be multiples of 8");
  ALDM_REQUIRE(d.ld_t >= d.Nk, ALDM_E_SHAPE, "attention: ld_t=%d < Nk=%d", d.ld_t, d.Nk);
  ALDM_REQUIRE(aligned16(d.q_hi) && aligned16(d.q_lo) && aligned16(d.k_hi) && aligned16(d.k_lo) && aligned16(d.vt_hi) &&
                   aligned16(d.vt_lo) && aligned16(d.out_hi) && aligned16(d.out_lo),
               ALDM_E_ALIGN, "attention: pointers must be 16B aligned");
  ALDM_REQUIRE(d.heads <= 65535 && d.B <= 65535, ALDM_E_SHAPE, "attention: grid too large");
  if (d.impl == ALDM_GEMM_SIMT) {
    dim3 grid(cdiv(d.Nq, 128), d.heads, d.B);
    attention_simt_kernel<<<grid, 128, 0, st>>>(d);
  } else if (d.Nk <= 32 && !(getenv("ALDM_ATTN_SHORT") && getenv("ALDM_ATTN_SHORT")[0] == '0')) {
    dim3 grid(cdiv(d.Nq, 128), d.heads, d.B);
    if (d.Nk <= 8) ALDM_CHECK_CUDA(launch_pdl(attention_short_kernel<8>, grid, dim3(128), 0, st, d));
    else if (d.Nk <= 16) ALDM_CHECK_CUDA(launch_pdl(attention_short_kernel<16>, grid, dim3(128), 0, st, d));
    else ALDM_CHECK_CUDA(launch_pdl(attention_short_kernel<32>, grid, dim3(128), 0, st, d));
  } else {
    static int stages = 0;
    if (stages == 0) {
      ALDM_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::Cfg<2>::SMEM));
      ALDM_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::Cfg<3>::SMEM));
      // Both layouts are 97 KB (two CTAs per SM).  cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for either
      // (the 2-stage kernel measurably runs two per SM under ncu), so it is not consulted.  ALDM_ATTN_STAGES=2 = A/B switch.
      const char* e = getenv("ALDM_ATTN_STAGES");
      stages = (e && e[0] == '2') ? 2 : 3;
    }
    dim3 grid(cdiv(d.Nq, atc::QT), d.heads, d.B);
    if (stages == 3) ALDM_CHECK_CUDA(launch_pdl(attention_tc_kernel<3>, grid, dim3(192), atc::Cfg<3>::SMEM, st, d));
    else ALDM_CHECK_CUDA(launch_pdl(attention_tc_kernel<2>, grid, dim3(192), atc::Cfg<2>::SMEM, st, d));
  }
  ALDM_CHECK_CUDA(cudaGetLastError());
  return ALDM_OK;
}

// row softmax of x[rows, n] (x already scaled when scale == 1) -> planes [rows, n]; one block per row
__global__ void softmax_rows_kernel(const float* __restrict__ x, int n, float scale, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* xp = x + row * n;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, xp[i] * scale);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += expf(xp[i] * scale - mx);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
  const float inv = 1.0f / s;
  for (int i = threadIdx.x * 2; i < n; i += blockDim.x * 2) {
    const float a = expf(xp[i] * scale - mx) * inv;
    const float b = (i + 1 < n) ? expf(xp[i + 1] * scale - mx) * inv : 0.f;
    uint32_t h, l;
    split2(a, b, h, l);
    if (i + 1 < n) {
      *reinterpret_cast<uint32_t*>(hi + row * n + i) = h;
      *reinterpret_cast<uint32_t*>(lo + row * n + i) = l;
    } else {
      hi[row * n + i] = __float2bfloat16_rn(a);
      lo[row * n + i] = __float2bfloat16_rn(a - __bfloat162float(__float2bfloat16_rn(a)));
    }
  }
}

}  // namespace aldm

extern "C" int aldm_attention(const aldm_attn_desc* d, void* stream) {
  if (!d) { aldm::set_error("aldm_attention: null desc"); return ALDM_E_ARG; }
  return aldm::attention_launch(*d, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int aldm_softmax_rows(const float* x, int32_t rows, int32_t n, float scale, void* out_hi, void* out_lo,
                                 void* stream) {
  using namespace aldm;
  ALDM_REQUIRE(x && out_hi && out_lo && rows > 0 && n > 0, ALDM_E_ARG, "softmax_rows: bad arguments");
  ALDM_REQUIRE(n % 2 == 0, ALDM_E_UNSUPPORTED, "softmax_rows: n must be even");
  softmax_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, n, scale, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo));
  ALDM_CHECK_CUDA(cudaGetLastError());
  return ALDM_OK;
}
