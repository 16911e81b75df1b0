// Shared device helpers for the sm_100a kernels: PTX wrappers (mbarrier, cp.async, bulk copy,
// tcgen05 / TMEM), bf16 hi/lo splitting, error plumbing.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/aldm_b200.h"

namespace aldm {

void set_error(const char* fmt, ...);

#define ALDM_CHECK_CUDA(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::aldm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ALDM_E_CUDA;                                                            \
    }                                                                                \
  } while (0)

#define ALDM_REQUIRE(cond, code, ...)                                                \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      ::aldm::set_error(__VA_ARGS__);                                                \
      return (code);                                                                 \
    }                                                                                \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// Every kernel of the per-step programs is launched with programmatic stream serialisation: it may
// start (block scheduling, barrier init, TMEM allocation, index set-up) while its predecessor drains,
// and calls pdl_wait() before its first global-memory access, which blocks until the predecessor has
// completed and flushed.  Because every such kernel waits, completion stays transitive along the
// stream (kernel k+2 cannot pass its wait before kernel k is done).  pdl_launch() is issued LATE (last MMA
// issued / last loads done): triggering at kernel entry made small dependent blocks co-resident with the
// primary for its whole run time and measurably slowed it (+1.5 ms per DDIM step; split-K GEMM + reduction
// pair +12 us), so the early start is limited to the primary's tail.  It must also come AFTER any TMEM
// allocation (a dependent CTA that grabbed TMEM first could starve a still-unallocated primary CTA).
// ALDM_PDL=0 in the environment disables the launch attribute (the device instructions are then no-ops).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------
// fp16 operand planes.  hi = fp16(x) carries 11 significand bits; lo = fp16(x - hi) adds 11 more (|x - hi - lo| <=
// 2^-22 |x| while lo stays normal, <= 2^-25 absolute once it is subnormal).  Two-plane operands feed three tensor-core
// passes (hi*hi + hi*lo + lo*hi), single-plane activations two (hi*w_hi + hi*w_lo).  Inputs saturate at the fp16 range
// (+-65504) so that x - hi can never be inf - inf; activations behind a normalisation / gating stay far inside it.
// ------------------------------------------------------------------------------------------
typedef __half aldm_plane_t;
__device__ __forceinline__ float sat_f16(float x) { return fminf(fmaxf(x, -65504.0f), 65504.0f); }
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = sat_f16(a);
  b = sat_f16(b);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack2_hi(float a, float b) {       // hi plane only
  const __half2 h = __floats2half2_rn(sat_f16(a), sat_f16(b));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ float plane_to_f(aldm_plane_t v) { return __half2float(v); }
// scalar element store: hi (and lo when the operand has a second plane)
__device__ __forceinline__ void store_split1(aldm_plane_t* hp, aldm_plane_t* lp, long long i, float v) {
  v = sat_f16(v);
  const __half h = __float2half_rn(v);
  hp[i] = h;
  if (lp) lp[i] = __float2half_rn(v - __half2float(h));
}

// 8 consecutive floats -> one 16-byte chunk of hi and one of lo
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  split2(v[0], v[1], hi.x, lo.x);
  split2(v[2], v[3], hi.y, lo.y);
  split2(v[4], v[5], hi.z, lo.z);
  split2(v[6], v[7], hi.w, lo.w);
}
// same without the fp16-range clamp, for values known to lie in [0, 1] (softmax probabilities): the clamp is two FMNMX per
// element, a fifth of the attention softmax loop's instructions
__device__ __forceinline__ uint4 pack8_hi_unit(const float* v) {
  uint4 r;
  __half2 h;
  h = __floats2half2_rn(v[0], v[1]); r.x = *reinterpret_cast<const uint32_t*>(&h);
  h = __floats2half2_rn(v[2], v[3]); r.y = *reinterpret_cast<const uint32_t*>(&h);
  h = __floats2half2_rn(v[4], v[5]); r.z = *reinterpret_cast<const uint32_t*>(&h);
  h = __floats2half2_rn(v[6], v[7]); r.w = *reinterpret_cast<const uint32_t*>(&h);
  return r;
}
__device__ __forceinline__ uint4 pack8_hi(const float* v) {
  return make_uint4(pack2_hi(v[0], v[1]), pack2_hi(v[2], v[3]), pack2_hi(v[4], v[5]), pack2_hi(v[6], v[7]));
}

// x * sigmoid(x) with approximate reciprocal (the IEEE division expands to ~20 instructions with a guarded slow path;
// the GroupNorm+SiLU apply kernel is issue-bound on it).  ~2 ulp, far below the 2^-17 operand split that follows.
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// erf GELU as F.gelu default (attention.py:44).  erf via Abramowitz-Stegun 7.1.26 (|err| < 5e-7 in fp32,
// i.e. at fp32 round-off of the GELU output): 1 RCP + 1 EX2 + 8 FMA-class instructions instead of the
// ~40-instruction erff -- the GEGLU epilogue is issue-bound (profiles/r01_gemm_timeline_tc3.txt).
__device__ __forceinline__ float gelu_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;      // rcp.approx (1 MUFU, ~1 ulp): __frcp_rn expands to MUFU + Newton step + a guarded slow-path CALL per element
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-p, e, 1.0f);
  return 0.5f * x + 0.5f * fabsf(x) * erf_abs;     // 0.5 x (1 + sign(x) erf(|x|/sqrt2))
}

// v[i] *= gelu(g[i]) for NB values at once, written stage by stage so that NB independent dependency chains are in
// flight: the GEGLU epilogue runs two warps per scheduler, and with the elements evaluated one or two at a time
// (what the compiler produced from the scalar form under the 128-register cap) the ~75-cycle chain of each element
// (two MUFU round trips) was fully exposed: 3,900 cycles per 32 x 32 chunk in the timeline, 7,000 per tile.
template <int NB>
__device__ __forceinline__ void geglu_mul(float* __restrict__ v, const float* __restrict__ g) {
  float z[NB], t[NB], e[NB], p[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) z[j] = fabsf(g[j]) * 0.70710678118654752440f;
#pragma unroll
  for (int j = 0; j < NB; ++j) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(fmaf(0.3275911f, z[j], 1.0f)));
#pragma unroll
  for (int j = 0; j < NB; ++j) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(-z[j] * z[j] * 1.4426950408889634f));
#pragma unroll
  for (int j = 0; j < NB; ++j) p[j] = fmaf(1.061405429f, t[j], -1.453152027f);
#pragma unroll
  for (int j = 0; j < NB; ++j) p[j] = fmaf(p[j], t[j], 1.421413741f);
#pragma unroll
  for (int j = 0; j < NB; ++j) p[j] = fmaf(p[j], t[j], -0.284496736f);
#pragma unroll
  for (int j = 0; j < NB; ++j) p[j] = fmaf(p[j], t[j], 0.254829592f);
#pragma unroll
  for (int j = 0; j < NB; ++j) p[j] *= t[j] * e[j];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const float erf_abs = 1.0f - p[j];
    v[j] *= fmaf(0.5f * fabsf(g[j]), erf_abs, 0.5f * g[j]);      // 0.5 g (1 + sign(g) erf(|g| / sqrt 2))
  }
}

// ------------------------------------------------------------------------------------------
// shared-memory address + mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must become a CUDA error, never a hung GPU box.  The timer is only
// consulted every 4096 failed probes (try_wait itself suspends the thread in hardware), so the
// hot path is a bare try_wait loop.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 4095u) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {   // 4 s
        printf("aldm: mbarrier wait timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
               blockIdx.z, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

// non-blocking probe (mbarrier.test_wait never suspends the thread)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// pure spin variant for latency-critical single-thread waits
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if (++spins > (1u << 28)) { printf("aldm: mbarrier spin timeout\n"); __trap(); }
  }
}

// ------------------------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte with zero fill, completion signalled on an mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// same with an immediate destination offset (one base register for the eight rows a producer thread copies per k-block)
template <int OFF>
__device__ __forceinline__ void cp_async_16_off(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0 + %3], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes), "n"(OFF) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// TMA bulk copy global -> shared, completion (bytes) on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 2-D tiled tensor-map load (TMA): box lands in shared memory in the map's swizzle mode; completes on `mbar` (tx bytes)
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tensor_map, int c0, int c1, uint32_t mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tensor_map)), "r"(mbar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One elected lane of a fully converged warp.  The single-thread roles (tcgen05.mma / commit issue, TMA bulk copies) must
// be entered through THIS predicate, not `lane == 0`: the instructions take their operands from uniform registers, and
// under a lane-id predicate the compiler cannot prove uniformity, so it wraps EVERY such instruction in an
// ELECT / BRA.U.ANY loop -- measured at ~95 cycles per tcgen05.mma (csrc/microbench.cu) against 64 cycles of tensor
// work for a 128 x 128 x 16 step, i.e. the issue loop, not the tensor core, paced every GEMM and the attention kernel.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, 16-bit float operands (format in the instruction descriptor), fp32 accumulate, cta_group::1
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (base_lane + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {      // one column: thread i gets lane (base_lane + i)
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// Accumulator read of the GEMM: the tile lives in TMEM as two column blocks DIST apart -- [A W_hi^T (+ A_lo W_hi^T) | A W_lo^T],
// written by ONE N = 2 BN tcgen05.mma per K step (gemm.cu) -- and the value is their sum.  The second block is read in two
// 16-column pieces so that at most 48 registers are live.  Includes the tcgen05.wait::ld.
template <int DIST, bool STACKED>
__device__ __forceinline__ void acc_ld32(uint32_t taddr, uint32_t* r) {
  if (!STACKED) {
    tmem_ld32(taddr, r);
    tmem_ld_wait();
    return;
  }
  uint32_t t[16];
  tmem_ld32(taddr, r);
  tmem_ld16(taddr + DIST, t);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(t[i]));
  tmem_ld16(taddr + DIST + 16, t);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) r[16 + i] = __float_as_uint(__uint_as_float(r[16 + i]) + __uint_as_float(t[i]));
}

// K-major, 128-byte-swizzled operand tile descriptor (rows of 128 B, 8-row groups 1024 B apart).
// Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
// Instruction descriptor (InstrDescriptor): c_format F32=1 [4,6), a/b_format F16=0 [7,10)/[10,13) (BF16 would be 1),
// a/b major K=0, n_dim = N>>3 [17,23), m_dim = M>>4 [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace aldm
