// K1/K2/K7/K8: implicit-GEMM convolution / linear layer on tcgen05 tensor cores.
//
//   out[row(m), n] = epilogue( sum_k A[m,k] * W[n,k] )        (aldm_gemm_desc, include/aldm_b200.h)
//
// Precision: split-fp16 operands, fp32 accumulation in TMEM.  Weights are packed as two fp16 planes (w ~= hi + lo,
// 22 significand bits).  Activations arrive as fp16 planes written by the prep kernels / producer epilogues:
//   * two planes (a_lo != NULL): the products hi*hi + hi*lo_w + lo*hi_w, ~2^-22 operand precision, issued as TWO kind::f16
//     instructions per K step (A_hi against the stage's [W_hi ; W_lo] as one N = 2 BN operand, A_lo against W_hi) -- the
//     convolutions, where the 200-step waveform budget goes (scripts/precision_study.py);
//   * one plane (a_lo == NULL): two instructions (hi*lo_w, hi*hi_w) and half the A bytes -- the token-side linear layers.
// (plain single-pass bf16 / TF32-class rounding of BOTH operands misses or crowds the 1e-3 waveform
// tolerance; rounding only the activations to 11 bits costs 2e-4 at 200 steps.)
//
// Structure of one CTA (persistent: 448 threads, one CTA per SM looping over 128 x BN output tiles / split-K slices; details
// and the measurements behind them at gemm_tc3_kernel below and in DESIGN.md section 4):
//   warps 0-3   A producers: gather 16-byte chunks (8 channels of one tap of one pixel) with cp.async + zero fill into
//               128B-swizzled K-major tiles; completion is signalled on the stage's mbarrier (cp.async.mbarrier.arrive).
//   warp 4      B producer: one ELECTED lane (elect.sync) issues a TMA bulk copy (cp.async.bulk) of the host-packed,
//               pre-swizzled weight tile image (hi|lo) per stage.
//   warp 5      allocates TMEM; one elected lane issues tcgen05.mma into a double-buffered accumulator and commits stages.
//   warps 6-13  epilogue out of TMEM, overlapping the next tile's main loop.
#include <stdlib.h>

#include "common.cuh"

namespace aldm {

// ------------------------------------------------------------------------------------------
// row decoding + epilogue shared by the tensor-core kernel, the SIMT checker and split-K
// ------------------------------------------------------------------------------------------
struct RowInfo {
  int m, b, oh, ow;
  bool valid;
  long long orow;
};

__device__ __forceinline__ RowInfo decode_row(const aldm_gemm_desc& d, int m, int M) {
  RowInfo r;
  r.m = m;
  r.valid = m < M;
  int mm = r.valid ? m : 0;
  r.ow = mm % d.OW;
  int t = mm / d.OW;
  r.oh = t % d.OH;
  r.b = t / d.OH;
  r.orow = ((long long)r.b * d.OHF + (long long)r.oh * d.osy + d.ooy) * d.OWF + r.ow;
  return r;
}

// v[0..cnt) are post-activation values for output columns [n0, n0+cnt) of row r.
__device__ __forceinline__ void epi_finish(const aldm_gemm_desc& d, const RowInfo& r, int n0, int cnt, float* v,
                                           int n_out) {
  if (!r.valid) return;
  if (n0 >= n_out) return;
  if (n0 + cnt > n_out) cnt = n_out - n0;
  if (d.res) {
    const float* rp = d.res + r.orow * d.ld_res + n0;
    for (int i = 0; i < cnt; ++i) v[i] += __ldg(rp + i);
  }
  if (d.alpha != 1.0f)
    for (int i = 0; i < cnt; ++i) v[i] *= d.alpha;
  if (d.out_mode == ALDM_OUT_F32) {
    float* op = d.out + r.orow * d.ldo + n0;
    if (d.accumulate)
      for (int i = 0; i < cnt; ++i) v[i] += op[i];
    if (cnt == 32 && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
      for (int i = 0; i < cnt; ++i) op[i] = v[i];
    }
    if (d.out_hi) {   // dual output: the same values also as operand planes for the next GEMM
      aldm_plane_t* hp = reinterpret_cast<aldm_plane_t*>(d.out_hi) + r.orow * d.ldo + n0;
      aldm_plane_t* lp = d.out_lo ? reinterpret_cast<aldm_plane_t*>(d.out_lo) + r.orow * d.ldo + n0 : nullptr;
      for (int i = 0; i < cnt; ++i) store_split1(hp, lp, i, v[i]);
    }
  } else if (d.out_mode == ALDM_OUT_QKV && n0 >= d.n_split) {
    // V projection: transposed planes [(b*Cv + c), ld_t] with the token index contiguous
    const int b = r.m / d.tok_per_batch, tok = r.m - b * d.tok_per_batch;
    const long long base = ((long long)b * (d.N - d.n_split) + (n0 - d.n_split)) * d.ld_t + tok;
    aldm_plane_t* hp = reinterpret_cast<aldm_plane_t*>(d.out2_hi) + base;
    aldm_plane_t* lp = d.out2_lo ? reinterpret_cast<aldm_plane_t*>(d.out2_lo) + base : nullptr;
    for (int i = 0; i < cnt; ++i) store_split1(hp, lp, (long long)i * d.ld_t, v[i]);
    // keys in [tok_per_batch, ld_t) are padding the attention kernel multiplies by P = 0: they must
    // be finite (stale workspace bytes reinterpreted as fp16 could be NaN), so the last token zeroes them
    if (tok == d.tok_per_batch - 1) {
      for (int t = 1; tok + t < d.ld_t; ++t)
        for (int i = 0; i < cnt; ++i) store_split1(hp, lp, (long long)i * d.ld_t + t, 0.f);
    }
  } else if (d.out_mode == ALDM_OUT_PLANES || d.out_mode == ALDM_OUT_QKV) {
    aldm_plane_t* hp = reinterpret_cast<aldm_plane_t*>(d.out_hi) + r.orow * d.ldo + n0;
    aldm_plane_t* lp = d.out_lo ? reinterpret_cast<aldm_plane_t*>(d.out_lo) + r.orow * d.ldo + n0 : nullptr;
    if (cnt == 32 && ((reinterpret_cast<uintptr_t>(hp) & 15u) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 h, l;
        split8(v + i, h, l);
        *reinterpret_cast<uint4*>(hp + i) = h;
        if (lp) *reinterpret_cast<uint4*>(lp + i) = l;
      }
    } else {
      for (int i = 0; i < cnt; ++i) store_split1(hp, lp, i, v[i]);
    }
  } else {  // NCHW
    for (int i = 0; i < cnt; ++i)
      d.out[(((long long)r.b * d.N + (n0 + i)) * d.OH + r.oh) * d.OW + r.ow] = v[i];
  }
}

// Bias / rowvec / activation for one 32-column chunk whose packed column base is pc0.
// For GEGLU `g` holds the gate chunk (packed columns pc0 + bn/2 ...).
__device__ __forceinline__ void add_vec32(float* v, const float* __restrict__ p) {
  if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
      v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __ldg(p + i);
  }
}

__device__ __forceinline__ void epi_activate(const aldm_gemm_desc& d, const RowInfo& r, int pc0, float* v, float* g) {
  if (d.bias) {
    add_vec32(v, d.bias + pc0);
    if (d.act == ALDM_ACT_GEGLU) add_vec32(g, d.bias + pc0 + d.bn / 2);
  }
  if (d.rowvec) add_vec32(v, d.rowvec + (long long)r.b * d.ld_rowvec + pc0);
  if (d.act == ALDM_ACT_GEGLU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= gelu_f(g[i]);
  } else if (d.act == ALDM_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
  } else if (d.act == ALDM_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
  }
}

// Coalesced finish: the 32x32 chunk (lane == row) is transposed through a per-warp shared-memory
// staging tile (stride 33: conflict-free both ways) so that 8 lanes cover one 128-byte output row
// segment: residual loads and output stores are full-line float4 transactions instead of 32
// scattered 16-byte pieces.  Handles fp32 and operand-plane outputs; returns false if the layout
// does not allow it (caller falls back to the row-owner path).
__device__ __forceinline__ bool epi_coalescable(const aldm_gemm_desc& d, int n_out) {
  if (n_out % 4 != 0 || d.ldo % 4 != 0) return false;
  if (d.res && d.ld_res % 4 != 0) return false;
  return d.out_mode == ALDM_OUT_F32 || d.out_mode == ALDM_OUT_PLANES || d.out_mode == ALDM_OUT_QKV;
}

// Per-tile row bookkeeping of the coalesced path: lane (rs = lane>>3, c4 = (lane&7)*4) handles rows
// it*4+rs, it = 0..7; orow[it] / valid bits are fetched once per tile from the row-owner lanes.
struct CoRows {
  long long orow[8];
  unsigned vmask;
};

__device__ __forceinline__ CoRows co_rows(const RowInfo& r, int lane) {
  CoRows cr;
  cr.vmask = __ballot_sync(0xffffffffu, r.valid);
  const int rs = lane >> 3;
#pragma unroll
  for (int it = 0; it < 8; ++it) cr.orow[it] = __shfl_sync(0xffffffffu, r.orow, it * 4 + rs);
  return cr;
}

// issue the residual loads of one 32-column chunk (they are consumed after the TMEM read + transpose,
// and the next chunk's loads are issued before the current chunk is processed: latency hidden)
__device__ __forceinline__ void co_load_res(const aldm_gemm_desc& d, const CoRows& cr, int n0, int n_out, int lane, float4 (&rv)[8]) {
  const int rs = lane >> 3, n = n0 + (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const bool ok = ((cr.vmask >> (it * 4 + rs)) & 1u) && n < n_out;
    rv[it] = ok ? __ldg(reinterpret_cast<const float4*>(d.res + cr.orow[it] * d.ld_res + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ void epi_finish_coalesced(const aldm_gemm_desc& d, const CoRows& cr, int n0, const float* v,
                                                     int n_out, float* stg, int lane, const float4 (&rv)[8], bool has_rv) {
#pragma unroll
  for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = v[i];
  __syncwarp();
  const int rs = lane >> 3, c4 = (lane & 7) * 4;
  const int n = n0 + c4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rs;
    if (!((cr.vmask >> rr) & 1u) || n >= n_out) continue;
    const long long orow = cr.orow[it];
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = stg[rr * 33 + c4 + i];
    if (has_rv) { x[0] += rv[it].x; x[1] += rv[it].y; x[2] += rv[it].z; x[3] += rv[it].w; }
    if (d.alpha != 1.0f) {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] *= d.alpha;
    }
    if (d.out_mode == ALDM_OUT_F32) {
      float4* op = reinterpret_cast<float4*>(d.out + orow * d.ldo + n);
      if (d.accumulate) {
        const float4 t = *op;
        x[0] += t.x; x[1] += t.y; x[2] += t.z; x[3] += t.w;
      }
      *op = make_float4(x[0], x[1], x[2], x[3]);
    }
    if (d.out_mode != ALDM_OUT_F32 || d.out_hi) {
      uint2 h, l;
      split2(x[0], x[1], h.x, l.x);
      split2(x[2], x[3], h.y, l.y);
      *reinterpret_cast<uint2*>(reinterpret_cast<aldm_plane_t*>(d.out_hi) + orow * d.ldo + n) = h;
      if (d.out_lo) *reinterpret_cast<uint2*>(reinterpret_cast<aldm_plane_t*>(d.out_lo) + orow * d.ldo + n) = l;
    }
  }
  __syncwarp();
}

// ---- compact coalesced finish (EPI_F32N / EPI_PLN / EPI_GEGLU) -----------------------------------
// Same idea as above with a quarter of the instructions: the staging tile is 32 rows x 128 bytes with the
// 16-byte chunk index XOR-swizzled by (row & 7), so both the row-owner writes and the transposed reads are
// conflict-free 128-bit accesses (8 STS.128 + 8 LDS.128 per lane instead of 32 + 32 scalar ones); bias is
// added after the transpose (one float4 per lane and chunk, fetched before the accumulator is ready), the
// output mode is a template parameter, and alpha / accumulate / rowvec are not supported (host-checked).
struct CoRows32 {       // host guarantees rows * max(ldo, ld_res) < 2^31 on this path
  int orow[8];
  unsigned vmask;
};

__device__ __forceinline__ CoRows32 co_rows32(const RowInfo& r, int lane) {
  CoRows32 cr;
  cr.vmask = __ballot_sync(0xffffffffu, r.valid);
  const int rs = lane >> 3;
  const int o = (int)r.orow;
#pragma unroll
  for (int it = 0; it < 8; ++it) cr.orow[it] = __shfl_sync(0xffffffffu, o, it * 4 + rs);
  return cr;
}

__device__ __forceinline__ void co_load_res32(const aldm_gemm_desc& d, const CoRows32& cr, int n0, int n_out, int lane, float4 (&rv)[8]) {
  const int rs = lane >> 3, n = n0 + (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const bool ok = ((cr.vmask >> (it * 4 + rs)) & 1u) && n < n_out;
    rv[it] = ok ? __ldg(reinterpret_cast<const float4*>(d.res + (cr.orow[it] * d.ld_res + n))) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ void stage_rows(uint8_t* stg, int lane, const float* v) {
  uint8_t* wr = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(wr + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
}

template <bool PLANES_ONLY, typename CR>
__device__ __forceinline__ void emit_rows(const aldm_gemm_desc& d, const CR& cr, int n0, int n_out, const uint8_t* stg,
                                          int lane, const float4 (&rv)[8], bool has_rv, float4 b4) {
  const int rs = lane >> 3, c8 = lane & 7;
  const int n = n0 + c8 * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + rs;
    float4 x = *reinterpret_cast<const float4*>(stg + rr * 128 + ((c8 ^ (rr & 7)) << 4));
    if (!((cr.vmask >> rr) & 1u) || n >= n_out) continue;
    x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
    if (has_rv) { x.x += rv[it].x; x.y += rv[it].y; x.z += rv[it].z; x.w += rv[it].w; }
    const auto o = cr.orow[it] * d.ldo + n;      // 32-bit on the compact path, 64-bit for GEGLU (CoRows)
    if (!PLANES_ONLY) *reinterpret_cast<float4*>(d.out + o) = x;
    if (PLANES_ONLY || d.out_hi) {
      uint2 h, l;
      split2(x.x, x.y, h.x, l.x);
      split2(x.z, x.w, h.y, l.y);
      *reinterpret_cast<uint2*>(reinterpret_cast<aldm_plane_t*>(d.out_hi) + o) = h;
      if (d.out_lo) *reinterpret_cast<uint2*>(reinterpret_cast<aldm_plane_t*>(d.out_lo) + o) = l;
    }
  }
  __syncwarp();
}

// ---- full-line finish for single-plane fp16 outputs (EPI_PLN, EPI_GEGLU -> planes) ------------------------
// A warp's 32 x 32 chunk is only 64 bytes per row in fp16: stored by itself it is a stream of half-line transactions, and the
// SM's store port moves one transaction per clock whatever its size (32 B/clk for full lines: profiles/r02_store_port_rate.txt;
// the GEGLU epilogue spent 2,500 of its 5,100 cycles per tile storing 16 KB).  The two warps that own the same TMEM lane quarter
// (chunk parity 0 / 1: columns [0,32) and [32,64) of one 64-column group) therefore assemble the 32 x 64 fp16 block in a shared
// tile (rows of 128 bytes, 16-byte chunks XOR-swizzled by the row), meet at a 64-thread named barrier, and each stores 16
// complete 128-byte rows: 4 STG.128 per warp instead of 8 STG.64.  Two tiles alternate, so one barrier per block suffices.
template <typename CR>
__device__ __forceinline__ void emit_pair_hi(const aldm_gemm_desc& d, const CR& cr, int n_pair0, uint8_t* tile, int lane, int half,
                                             int bar_id, const float* v) {
  uint8_t* wr = tile + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(wr + (((half * 4 + j) ^ sw) << 4)) = pack8_hi(v + 8 * j);
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  const int rs = lane >> 3, c8 = lane & 7;
  aldm_plane_t* out = reinterpret_cast<aldm_plane_t*>(d.out_hi);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = half * 16 + it * 4 + rs;
    const uint4 x = *reinterpret_cast<const uint4*>(tile + rr * 128 + ((c8 ^ (rr & 7)) << 4));
    const auto orow = half ? cr.orow[4 + it] : cr.orow[it];
    if ((cr.vmask >> rr) & 1u) *reinterpret_cast<uint4*>(out + (orow * d.ldo + n_pair0 + c8 * 8)) = x;
  }
}

// ------------------------------------------------------------------------------------------
// tensor-core kernel
// ------------------------------------------------------------------------------------------
template <int BN, int AP>
struct Tc3Cfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;                       // fp16 elements = 128 bytes per row
  static constexpr int A_BYTES = BM * 128;            // one plane
  static constexpr int B_BYTES = BN * 128;            // one plane
  static constexpr int STAGE_BYTES = AP * A_BYTES + 2 * B_BYTES;      // [a_hi | a_lo (AP == 2)] [b_hi | b_lo]
  static constexpr int B_OFF = AP * A_BYTES;
  static constexpr int STG_BYTES = 8 * 32 * 33 * 4;   // one 32x33 fp32 transpose tile per epilogue warp
  static constexpr int SMEM_MAX = 227 * 1024;
  static constexpr int FIT = (SMEM_MAX - 1024 - 256 - STG_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 4 ? 4 : FIT;    // BN=128: 3 (AP=2, 64 KB stages) / 4 (AP=1, 48 KB)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STG_BYTES;
  // ST ("stacked"): the B stage [W_hi rows | W_lo rows] is issued as ONE N = 2 BN operand, so the accumulator is two column
  // blocks [A W_hi^T (+ A_lo W_hi^T) | A W_lo^T] that the epilogue adds.  Used for the two-plane (convolution) operands, whose
  // long K loops are tensor-bound: 2 instead of 3 instructions per K step (measured 273 vs 314 cycles, csrc/microbench.cu).  The
  // single-plane token-side GEMMs keep the plain form: their short K loops are epilogue-bound and the second TMEM read of the
  // stacked form made them slower (lin_k256_n768_qkv 34.6 -> 44.6 us), while their MMA time hides under the epilogue anyway.
  static constexpr bool ST = AP == 2;
  static constexpr int TMEM_COLS = ST ? 2 * BN : BN;  // one accumulator (power of two >= 32)
  static_assert(STAGES >= 2, "pipeline needs two stages");
};

// Debug timeline (profiling aid, dbg bit 128): CTA 0 records clock64() at pipeline events.
// layout: [role 0..3][iteration 0..255][phase 0..1]
__device__ long long g_timeline[4 * 256 * 2];
#define ALDM_TL(role, i, ph)                                                      \
  do {                                                                            \
    if ((dbg & 128) && blockIdx.x == 0 && (i) < 256) g_timeline[((role) * 256 + (i)) * 2 + (ph)] = clock64(); \
  } while (0)

// ------------------------------------------------------------------------------------------
// persistent variant (default): one CTA per SM loops over output tiles; 448 threads =
//   warps 0-3 A producers | warp 4 B (TMA bulk) | warp 5 MMA | warps 6-13 epilogue (two per TMEM lane
//   quarter, alternating 32-column chunks).
// Two TMEM accumulators (2 x BN columns) let the epilogue of tile i drain while the producers and the
// tensor core already work on tile i+1; barrier init / TMEM alloc are paid once per CTA.
// What the measured timeline (profiles/r01_gemm_timeline.txt) showed and this version fixes:
//   * each role runs ONE warp per SM sub-partition, so dependent-instruction latency is fully exposed:
//     the producers' per-stage address arithmetic (an integer division and eight 64-bit index chains)
//     took ~1450 cycles against 768 cycles of MMA work.  Row bases and per-row tap-validity masks are
//     now computed once per tile; a stage costs one add + one bit test per row.
//   * the epilogue took ~26k cycles per tile with four warps and a generic (all modes inlined, 200 KB
//     of SASS) body.  EPI selects a specialised body at compile time (0: linear/conv with optional
//     bias, row vector, residual, dual/QKV plane outputs; 1: GEGLU; 2: everything else), and eight
//     warps share the drain.
// Tile order: split-K slice fastest, then n-tile, then m-tile, so CTAs running concurrently share the
// same activation rows in L2.
// ------------------------------------------------------------------------------------------
enum { EPI_FAST = 0, EPI_GEGLU = 1, EPI_GENERIC = 2, EPI_F32N = 3, EPI_PLN = 4 };

// Division by a launch-time constant as multiply + shift (n < 2^31, d < 2^31): q = (n * M) >> (32 + l),
// M = floor(2^(32+l) / d) + 1, l = ceil(log2 d).  The persistent roles run one warp per scheduler, so a
// hardware-emulated integer division (~100 dependent instructions) per row per tile stalls the pipeline.
struct FastDiv {
  unsigned long long M;
  int sh;
  int d;
  __device__ __forceinline__ int div(int n) const {
    return (int)(((unsigned long long)(unsigned)n * M) >> sh);   // n >= 0
  }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * d; }
};
static FastDiv make_fastdiv(int d) {
  FastDiv f;
  int l = 0;
  while ((1ll << l) < d) ++l;
  f.M = (unsigned long long)((((unsigned __int128)1) << (32 + l)) / (unsigned)d) + 1ull;
  f.sh = 32 + l;
  f.d = d;
  if (d == 1) { f.M = 1ull << 32; f.sh = 32; }
  return f;
}
struct Tc3Divs {
  FastDiv ow, oh, cp, tn, bmod;
  int plain;      // 1x1 tap, unit stride, no upsample / batch-modulo: input row == output row (linear layers)
};

// AP = number of A planes (2: hi + lo, three UMMAs per K step; 1: hi only, two UMMAs and half the A bytes).
template <int BN, int EPI, int AP>
__global__ void __launch_bounds__(448, 1) gemm_tc3_kernel(const __grid_constant__ aldm_gemm_desc d, int tiles_m, int tiles_n,
                                                           const __grid_constant__ Tc3Divs fd) {
  using C = Tc3Cfg<BN, AP>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int M = d.B * d.OH * d.OW;
  const int nkb_total = d.Kpad / C::BK;
  const int total = tiles_m * tiles_n * d.splitk;
  const int dbg = d.impl >> 8;     // profiling aids (scripts/prof_ops.py --dbg): 1 skip A, 2 skip B, 4 skip MMA, 8 skip epilogue, 128 timeline

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 128 + 1);   // 128 cp.async producers (completion-triggered arrivals) + B expect_tx
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);        // tcgen05.commit after the tile's last k-block
      mbar_init(tempty_bar(a), 8);       // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 2 * C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Everything above overlapped the predecessor's tail; nothing below may touch its outputs before the wait.
  // The weight stream is the exception (ALDM_GEMM_STATIC_B): warp 4 starts filling the pipeline right away.
  // The A producers wait inside their branch, after the (memory-free) row decode of their first tile.
  if (warp > 4 || (warp == 4 && !(d.impl & ALDM_GEMM_STATIC_B))) pdl_wait();

  auto tile_coords = [&](int id, int& mt, int& nt, int& z, int& kb0, int& nkb) {
    int r = id;
    z = 0; kb0 = 0; nkb = nkb_total;
    if (d.splitk > 1) {
      z = id % d.splitk;
      r = id / d.splitk;
      kb0 = (int)(((long long)z * nkb_total) / d.splitk);
      nkb = (int)(((long long)(z + 1) * nkb_total) / d.splitk) - kb0;
    }
    fd.tn.divmod(r, mt, nt);
  };

  if (warp < 4) {
    // ===================== A producers =====================
    const int j = tid & 7;                 // 16-byte chunk (8 channels) inside the 64-wide K block
    const int rbase = tid >> 3;            // rows rbase + 16*i
    const uint32_t swz = (uint32_t)((j ^ (rbase & 7)) << 4);
    const int Hs = d.H >> d.up, Ws = d.W >> d.up;
    const aldm_plane_t* ahi = reinterpret_cast<const aldm_plane_t*>(d.a_hi);
    const aldm_plane_t* alo = reinterpret_cast<const aldm_plane_t*>(d.a_lo);
    uint32_t cnt = 0;
    int last_mt = -1;
    int rowoff[8];            // element offset of tap (0,0) / channel 0 of each row (valid rows only)
    uint32_t tapmask[8];      // bit (t + 4) set <=> tap t of this row is inside the input (pre-shifted: (mask >> t) & 16 = bytes to copy)
    int ih0[8], iw0[8], pbh[8];   // only used by the nearest-upsample (up = 1) slow path
    // Row decode of one M tile.  A single warp per scheduler runs this dependent integer chain at ~1 instruction
    // per 6-8 cycles, and the timeline showed ~5,000 idle tensor-core cycles at every tile boundary of the K = 256
    // linear layers because of it: linear layers take the trivial branch, and the first tile is decoded before the
    // programmatic-dependency wait (under the previous kernel's tail).
    auto decode_rows = [&](int mt) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = mt * C::BM + rbase + 16 * i;
        uint32_t msk = 0;
        int off = 0;
        ih0[i] = 0; iw0[i] = 0; pbh[i] = -1;
        if (fd.plain) {
          if (m < M) { msk = 16u; off = m * d.Cp; }
        } else if (m < M) {
          int t, ow, b, oh;
          fd.ow.divmod(m, t, ow);
          fd.oh.divmod(t, b, oh);
          const int y0 = oh * d.sy, x0 = ow * d.sx;
          int bs = b;
          if (d.bmod > 0) { int q; fd.bmod.divmod(b, q, bs); }
          const int pb = bs * Hs;
          ih0[i] = y0; iw0[i] = x0; pbh[i] = pb;
          for (int tp = 0; tp < d.ntaps; ++tp) {
            const int ih = y0 + d.dy[tp], iw = x0 + d.dx[tp];
            if (ih >= 0 && ih < d.H && iw >= 0 && iw < d.W) msk |= 16u << tp;
          }
          off = ((pb + y0) * Ws + x0) * d.Cp;
        }
        tapmask[i] = msk;
        rowoff[i] = off;
      }
    };
    if ((int)blockIdx.x < total) {
      int mt, nt, z, kb0, nkb;
      tile_coords(blockIdx.x, mt, nt, z, kb0, nkb);
      decode_rows(mt);
      last_mt = mt;
    }
    pdl_wait();
    for (int id = blockIdx.x; id < total; id += gridDim.x) {
      int mt, nt, z, kb0, nkb;
      tile_coords(id, mt, nt, z, kb0, nkb);
      if (mt != last_mt) {
        last_mt = mt;
        decode_rows(mt);
      }
      // (tap, c) of this thread's 8-channel chunk at the first k-block of the tile, then advanced by 64 per block
      const int k = kb0 * C::BK + j * 8;
      int tap, c;
      fd.cp.divmod(k, tap, c);
      for (int it = 0; it < nkb; ++it, ++cnt) {
        const int s = cnt % C::STAGES;
        mbar_wait(empty_bar(s), ((cnt / C::STAGES) & 1) ^ 1);
        if (tid == 0) ALDM_TL(0, cnt, 0);
        const bool kvalid = tap < d.ntaps;
        const int tp = kvalid ? tap : 0;
        const uint32_t sa = base + s * C::STAGE_BYTES + swz;
        // Per 16-byte copy the producers now issue 3 (linear) / 4 (conv) instructions instead of ~9: the destination is one
        // register + an immediate, the byte count (16 or 0 = zero fill) is a shift + mask of the per-row tap mask, and the
        // source is one 64-bit add on a per-k-block base.  (The timeline showed the four producer warps -- one per scheduler,
        // every dependent instruction fully exposed -- pacing the whole pipeline at ~710 cycles per k-block.)  A source
        // address whose byte count is 0 is never dereferenced (cp.async zero-fill), so it needs no clamping.
        if (fd.plain) {
          const uint32_t kb = (kvalid && c < d.Cp) ? 16u : 0u;
          const aldm_plane_t* ab = ahi + (kb ? c : 0);
          const aldm_plane_t* abl = alo + (kb ? c : 0);
          const uint32_t sa2 = sa + (uint32_t)rbase * 128u;
          if (!(dbg & 1)) {
#define ALDM_A_ROW(i)                                                                            \
            {                                                                                      \
              const uint32_t nb = tapmask[i] & kb;                                                 \
              cp_async_16_off<(i) * 2048>(sa2, ab + rowoff[i], nb);                                \
              if (AP == 2) cp_async_16_off<(i) * 2048 + C::A_BYTES>(sa2, abl + rowoff[i], nb);     \
            }
            ALDM_A_ROW(0) ALDM_A_ROW(1) ALDM_A_ROW(2) ALDM_A_ROW(3) ALDM_A_ROW(4) ALDM_A_ROW(5) ALDM_A_ROW(6) ALDM_A_ROW(7)
#undef ALDM_A_ROW
          }
        } else if (d.up == 0) {
          const int tapoff = (d.dy[tp] * Ws + d.dx[tp]) * d.Cp + c;
          const aldm_plane_t* ab = ahi + tapoff;
          const aldm_plane_t* abl = alo + tapoff;
          const int sh = kvalid ? tp : 27;                   // tapmask holds the validity bits pre-shifted by 4: (mask >> tp) & 16
          const uint32_t sa2 = sa + (uint32_t)rbase * 128u;
          if (!(dbg & 1)) {
#define ALDM_A_ROW(i)                                                                            \
            {                                                                                      \
              const uint32_t nb = (tapmask[i] >> sh) & 16u;                                        \
              cp_async_16_off<(i) * 2048>(sa2, ab + rowoff[i], nb);                                \
              if (AP == 2) cp_async_16_off<(i) * 2048 + C::A_BYTES>(sa2, abl + rowoff[i], nb);     \
            }
            ALDM_A_ROW(0) ALDM_A_ROW(1) ALDM_A_ROW(2) ALDM_A_ROW(3) ALDM_A_ROW(4) ALDM_A_ROW(5) ALDM_A_ROW(6) ALDM_A_ROW(7)
#undef ALDM_A_ROW
          }
        } else {      // nearest x2 upsample folded into the gather: source pixel = (ih >> 1, iw >> 1)
          const int dy = d.dy[tp], dx = d.dx[tp];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = kvalid && ((tapmask[i] >> (tp + 4)) & 1u);
            long long off = 0;
            if (ok) off = ((long long)(pbh[i] + ((ih0[i] + dy) >> 1)) * Ws + ((iw0[i] + dx) >> 1)) * d.Cp + c;
            const uint32_t dst = sa + (uint32_t)(rbase + 16 * i) * 128u;
            cp_async_16(dst, ahi + off, ok ? 16u : 0u);
            if (AP == 2) cp_async_16(dst + C::A_BYTES, alo + off, ok ? 16u : 0u);
          }
        }
        cp_async_mbar_arrive_noinc(full_bar(s));
        if (tid == 0) ALDM_TL(0, cnt, 1);
        c += C::BK;
        while (c >= d.Cp) { c -= d.Cp; ++tap; }
      }
    }
  } else if (warp == 4) {
    // ===================== B producer =====================
    if (elect_one()) {
      uint32_t cnt = 0;
      for (int id = blockIdx.x; id < total; id += gridDim.x) {
        int mt, nt, z, kb0, nkb;
        tile_coords(id, mt, nt, z, kb0, nkb);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(d.w_packed) + ((long long)nt * nkb_total + kb0) * (2 * C::B_BYTES);
        for (int it = 0; it < nkb; ++it, ++cnt) {
          const int s = cnt % C::STAGES;
          mbar_wait(empty_bar(s), ((cnt / C::STAGES) & 1) ^ 1);
          ALDM_TL(1, cnt, 0);
          if (dbg & 2) { mbar_arrive(full_bar(s)); continue; }
          mbar_arrive_expect_tx(full_bar(s), 2 * C::B_BYTES);
          bulk_g2s(base + s * C::STAGE_BYTES + C::B_OFF, wsrc + (long long)it * (2 * C::B_BYTES), 2 * C::B_BYTES,
                   full_bar(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(128, BN), idesc2 = umma_idesc_f16(128, 2 * BN);
      uint32_t cnt = 0, tl = 0;
      for (int id = blockIdx.x; id < total; id += gridDim.x, ++tl) {
        int mt, nt, z, kb0, nkb;
        tile_coords(id, mt, nt, z, kb0, nkb);
        const uint32_t acc = tl & 1;
        mbar_wait(tempty_bar(acc), ((tl >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * C::TMEM_COLS;
        for (int it = 0; it < nkb; ++it, ++cnt) {
          const int s = cnt % C::STAGES;
          mbar_wait(full_bar(s), (cnt / C::STAGES) & 1);
          ALDM_TL(2, cnt, 0);
          tc_fence_after();
          const uint32_t sa = base + s * C::STAGE_BYTES;
          const uint64_t da_hi = umma_desc_sw128(sa);
          const uint64_t da_lo = umma_desc_sw128(sa + C::A_BYTES);      // only used when AP == 2
          const uint64_t db_hi = umma_desc_sw128(sa + C::B_OFF);
          const uint64_t db_lo = umma_desc_sw128(sa + C::B_OFF + C::B_BYTES);      // only used by the plain (not stacked) form
          if (!(dbg & 4)) {
            // Measured (csrc/microbench.cu): a shared-memory-operand tcgen05.mma costs ~42 cycles + N / 2 -- the A fetch is
            // not overlapped -- so one N = 256 instruction (169 cycles) is cheaper than two N = 128 ones (2 x 105).
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {            // 4 x K=16 (32 bytes) inside the 128B swizzle row
              const uint64_t o = (uint64_t)(ks * 2);
              if (C::ST) {
                umma_f16(tacc, da_hi + o, db_hi + o, idesc2, (uint32_t)((it | ks) != 0));      // [A_hi W_hi^T | A_hi W_lo^T]
                umma_f16(tacc, da_lo + o, db_hi + o, idesc, 1);                                // + A_lo W_hi^T into the first block
              } else {
                umma_f16(tacc, da_hi + o, db_lo + o, idesc, (uint32_t)((it | ks) != 0));       // small term first
                umma_f16(tacc, da_hi + o, db_hi + o, idesc, 1);
              }
            }
          }
          umma_commit(empty_bar(s));
          ALDM_TL(2, cnt, 1);
        }
        umma_commit(tfull_bar(acc));
      }
      pdl_launch();     // all MMAs of this CTA are issued: let the next kernel's blocks be scheduled under the last epilogue
    }
    __syncwarp();
  } else {
    // ===================== epilogue: warps 6-13; TMEM lane quarter = warp % 4, chunk parity = (warp-6)/4 =====================
    const int lb = warp & 3;
    const int half = (warp - 6) >> 2;
    const int trow_in_tile = lb * 32 + lane;
    float* stg = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - raw)) + (warp - 6) * (32 * 33);
    uint32_t tl = 0, pair_cnt = 0;
    for (int id = blockIdx.x; id < total; id += gridDim.x, ++tl) {
      int mt, nt, z, kb0, nkb;
      tile_coords(id, mt, nt, z, kb0, nkb);
      const uint32_t acc = tl & 1;
      const int m = mt * C::BM + trow_in_tile;
      RowInfo r;
      {
        r.m = m; r.valid = m < M;
        const int mm = r.valid ? m : 0;
        int t;
        fd.ow.divmod(mm, t, r.ow);
        fd.oh.divmod(t, r.b, r.oh);
        r.orow = ((long long)r.b * d.OHF + (long long)r.oh * d.osy + d.ooy) * d.OWF + r.ow;
      }
      constexpr bool kCompact = EPI == EPI_F32N || EPI == EPI_PLN;
      CoRows cr;
      CoRows32 cr32;
      if (kCompact) cr32 = co_rows32(r, lane); else cr = co_rows(r, lane);
      // compact epilogues: bias and the first chunk's residual are fetched while the tile is still being accumulated
      constexpr int NCH = (BN + 63) / 64;       // 32-column chunks per epilogue warp
      constexpr int NRV = NCH > 1 ? 2 : 1;      // residual prefetch buffers (rotating)
      uint8_t* stg8 = reinterpret_cast<uint8_t*>(stg);
      float4 pb4[NCH], prv[NRV][8];
      const bool has_res = kCompact && d.res != nullptr;
      if (kCompact && d.splitk == 1) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int n = nt * BN + half * 32 + 64 * ch + (lane & 7) * 4;
          pb4[ch] = (d.bias && n < d.N) ? __ldg(reinterpret_cast<const float4*>(d.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // BOTH chunks' residuals are requested here, under the accumulator wait: issued after the first chunk was staged (the
        // previous version) the loads queued behind the chunk's stores in the LSU -- 1,000-2,000 cycles from "staged" to
        // "loads issued" in the timeline -- and the second chunk then waited for them.
#pragma unroll
        for (int ch = 0; ch < NRV; ++ch)
          if (has_res && half * 32 + 64 * ch < BN) co_load_res32(d, cr32, nt * BN + half * 32 + 64 * ch, d.N, lane, prv[ch]);
      }
      // full-line pair mode (emit_pair_hi): single fp16 plane out, no residual, whole 64-column groups
      const bool pair_pln = EPI == EPI_PLN && BN >= 64 && d.out_lo == nullptr && d.res == nullptr && d.splitk == 1 && d.N % 64 == 0 &&
                            d.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out_hi) & 15u) == 0;
      const bool pair_geglu = EPI == EPI_GEGLU && BN == 128 && d.out_mode == ALDM_OUT_PLANES && d.out_lo == nullptr && d.splitk == 1 &&
                              (d.N / 2) % 64 == 0 && d.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out_hi) & 15u) == 0;
      const bool pair_qk = EPI == EPI_FAST && BN >= 64 && d.out_mode == ALDM_OUT_QKV && d.out_lo == nullptr && d.res == nullptr &&
                           d.splitk == 1 && d.n_split % 64 == 0 && d.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out_hi) & 15u) == 0;
      uint8_t* pair_tiles = reinterpret_cast<uint8_t*>(smem_raw + (bar_base + 256 - raw)) + ((warp - 6) & 3) * (32 * 33 * 4);
      const int pair_bar = 1 + ((warp - 6) & 3);
      float pln_b[NCH > 0 ? NCH : 1];      // bias of this warp's columns (lane = column), distributed by shuffles in pair mode
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int n = nt * BN + half * 32 + 64 * ch + lane;
        pln_b[ch] = (pair_pln && d.bias && n < d.N) ? __ldg(d.bias + n) : 0.f;
      }
      float gb_v = 0.f, gb_g = 0.f;       // GEGLU bias of this warp's value / gate chunk (lane = column)
      if (EPI == EPI_GEGLU && d.bias && d.splitk == 1 && half * 32 < BN / 2) {
        gb_v = __ldg(d.bias + nt * BN + half * 32 + lane);
        gb_g = __ldg(d.bias + nt * BN + BN / 2 + half * 32 + lane);
      }
      if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8, 0);      // tile prologue (row decode, bias / residual prefetch) done
      mbar_wait(tfull_bar(acc), (tl >> 1) & 1);
      if (warp == 6 && lane == 0) ALDM_TL(3, tl, 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + acc * C::TMEM_COLS + ((uint32_t)(lb * 32) << 16);
      if (dbg & 8) {
        // skip
      } else if (d.splitk > 1) {
        // raw partial sums -> ws[z][m][n], coalesced through the staging tile
        const int Mpad = tiles_m * C::BM, Npad = tiles_n * BN;
        const int rs = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          acc_ld32<BN, C::ST>(trow + c0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = __uint_as_float(v[i]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rs;
            float* wp = d.ws + ((long long)z * Mpad + mt * C::BM + lb * 32 + rr) * Npad + nt * BN + c0 + c4;
            *reinterpret_cast<float4*>(wp) = make_float4(stg[rr * 33 + c4], stg[rr * 33 + c4 + 1], stg[rr * 33 + c4 + 2], stg[rr * 33 + c4 + 3]);
          }
          __syncwarp();
        }
      } else if (EPI == EPI_GEGLU) {
        // tile columns [0,BN/2) values, [BN/2,BN) gates; output width N/2, coalescable and without residual
        // (both checked on the host; the residual-free body keeps v/g in registers under the 128-register cap)
        const int n_out = d.N / 2;
        float4 rv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN / 2; c0 += 64) {
          const int n0 = nt * (BN / 2) + c0;
          uint32_t vr[32], gr[32];
          acc_ld32<BN, C::ST>(trow + c0, vr);
          acc_ld32<BN, C::ST>(trow + BN / 2 + c0, gr);
          if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 1, 0);
          float* v = reinterpret_cast<float*>(vr);
          float* g = reinterpret_cast<float*>(gr);
          // bias: fetched (one coalesced load per warp) BEFORE the accumulator wait, distributed by shuffles -- the broadcast
          // float4 loads it replaces sat between the TMEM read and the GELU with a full L2 round trip exposed
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] += __shfl_sync(0xffffffffu, gb_v, i);
            g[i] += __shfl_sync(0xffffffffu, gb_g, i);
          }
#pragma unroll      // full unroll: v/g must stay in registers (a partial unroll indexes them dynamically -> local memory)
          for (int i = 0; i < 32; i += 8) geglu_mul<8>(v + i, g + i);
          if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 1, 1);
          if (pair_geglu) {                         // the FF1 case: one fp16 plane for FF2, stored as full lines by the warp pair
            emit_pair_hi(d, cr, nt * (BN / 2), pair_tiles + ((pair_cnt++ & 1u) ? 4 * (32 * 33 * 4) : 0), lane, half, pair_bar, v);
            if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 2, 1);
          } else if (d.out_mode == ALDM_OUT_PLANES) {
            stage_rows(stg8, lane, v);
            if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 2, 0);
            emit_rows<true>(d, cr, n0, n_out, stg8, lane, rv, false, make_float4(0.f, 0.f, 0.f, 0.f));
            if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 2, 1);
          } else {
            epi_finish_coalesced(d, cr, n0, v, n_out, stg, lane, rv, false);
          }
        }
      } else if (kCompact && pair_pln) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c0 = half * 32 + 64 * ch;      // < BN: BN >= 64 in pair mode
          uint32_t vr[32];
          acc_ld32<BN, C::ST>(trow + c0, vr);
          float* v = reinterpret_cast<float*>(vr);
          if (d.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __shfl_sync(0xffffffffu, pln_b[ch], i);
          }
          emit_pair_hi(d, cr32, nt * BN + 64 * ch, pair_tiles + ((pair_cnt++ & 1u) ? 4 * (32 * 33 * 4) : 0), lane, half, pair_bar, v);
        }
      } else if (kCompact) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c0 = half * 32 + 64 * ch;
          if (c0 < BN) {
            const int n0 = nt * BN + c0;
            uint32_t vr[32];
            acc_ld32<BN, C::ST>(trow + c0, vr);
            if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 1 + 2 * ch, 0);
            stage_rows(stg8, lane, reinterpret_cast<const float*>(vr));
            if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 1 + 2 * ch, 1);
            if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 2 + 2 * ch, 0);
            emit_rows<EPI == EPI_PLN>(d, cr32, n0, d.N, stg8, lane, prv[ch % NRV], has_res, pb4[ch]);
            if (warp == 6 && lane == 0) ALDM_TL(3, 64 + tl * 8 + 2 + 2 * ch, 1);
          }
        }
      } else if (EPI == EPI_FAST) {
        // no activation; fp32 / planes / dual / QKV outputs, all through the coalesced path (host-checked)
        float4 rv[8];
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          const int n0 = nt * BN + c0;
          const bool vpart = d.out_mode == ALDM_OUT_QKV && n0 >= d.n_split;
          const bool pre = d.res != nullptr && !vpart;
          if (pre) co_load_res(d, cr, n0, d.N, lane, rv);
          uint32_t vr[32];
          acc_ld32<BN, C::ST>(trow + c0, vr);
          float* v = reinterpret_cast<float*>(vr);
          if (d.bias) add_vec32(v, d.bias + n0);
          if (d.rowvec) add_vec32(v, d.rowvec + (long long)r.b * d.ld_rowvec + n0);
          if (!vpart && pair_qk) {
            // Q | K planes (one fp16 plane, 64 bytes per row and chunk): full-line stores by the warp pair (emit_pair_hi)
            emit_pair_hi(d, cr, nt * BN + (c0 & ~63), pair_tiles + ((pair_cnt++ & 1u) ? 4 * (32 * 33 * 4) : 0), lane, half, pair_bar, v);
          } else if (!vpart) {
            epi_finish_coalesced(d, cr, n0, v, d.N, stg, lane, rv, pre);
          } else if (r.valid) {
            // V projection: transposed planes, lane == token -> consecutive lanes write consecutive bf16
            const int b = r.m / d.tok_per_batch, tok = r.m - b * d.tok_per_batch;
            const long long tb = ((long long)b * (d.N - d.n_split) + (n0 - d.n_split)) * d.ld_t + tok;
            aldm_plane_t* hp = reinterpret_cast<aldm_plane_t*>(d.out2_hi) + tb;
            aldm_plane_t* lp = d.out2_lo ? reinterpret_cast<aldm_plane_t*>(d.out2_lo) + tb : nullptr;
            const bool last = tok == d.tok_per_batch - 1;
#pragma unroll      // full unroll keeps v[] in registers
            for (int i = 0; i < 32; ++i) {
              if (n0 + i < d.N) store_split1(hp, lp, (long long)i * d.ld_t, v[i]);
            }
            if (last) {       // zero the padding keys [tok_per_batch, ld_t) (the attention kernel multiplies them by P = 0)
              for (int i = 0; i < 32 && n0 + i < d.N; ++i)
                for (int t = 1; tok + t < d.ld_t; ++t) store_split1(hp, lp, (long long)i * d.ld_t + t, 0.f);
            }
          }
        }
      } else {
        // generic: any activation / output mode, row-owner stores
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          if (d.act == ALDM_ACT_GEGLU) {
            if (c0 >= BN / 2) break;
            uint32_t vr[32], gr[32];
            acc_ld32<BN, C::ST>(trow + c0, vr);
            acc_ld32<BN, C::ST>(trow + BN / 2 + c0, gr);
            epi_activate(d, r, nt * BN + c0, reinterpret_cast<float*>(vr), reinterpret_cast<float*>(gr));
            epi_finish(d, r, nt * (BN / 2) + c0, 32, reinterpret_cast<float*>(vr), d.N / 2);
          } else {
            uint32_t vr[32];
            acc_ld32<BN, C::ST>(trow + c0, vr);
            epi_activate(d, r, nt * BN + c0, reinterpret_cast<float*>(vr), nullptr);
            epi_finish(d, r, nt * BN + c0, 32, reinterpret_cast<float*>(vr), d.N);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (warp == 6 && lane == 0) ALDM_TL(3, tl, 1);
    }
  }

  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// split-K reduction + epilogue
// ------------------------------------------------------------------------------------------
__global__ void splitk_epilogue_kernel(const __grid_constant__ aldm_gemm_desc d, int Mpad, int Npad) {
  pdl_wait();
  const int M = d.B * d.OH * d.OW;
  const int chunks_per_row = (d.act == ALDM_ACT_GEGLU) ? (Npad / d.bn) * (d.bn / 64) : Npad / 32;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * chunks_per_row) return;
  const int m = (int)(idx / chunks_per_row);
  const int ch = (int)(idx % chunks_per_row);
  const RowInfo r = decode_row(d, m, M);
  float v[32], g[32];
  int pc0, n0, n_out;
  if (d.act == ALDM_ACT_GEGLU) {
    const int per_tile = d.bn / 64;
    const int tile = ch / per_tile, sub = ch % per_tile;
    pc0 = tile * d.bn + sub * 32;
    n0 = tile * (d.bn / 2) + sub * 32;
    n_out = d.N / 2;
  } else {
    pc0 = ch * 32; n0 = pc0; n_out = d.N;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) { v[i] = 0.f; g[i] = 0.f; }
  for (int z = 0; z < d.splitk; ++z) {
    const float* wp = d.ws + ((long long)z * Mpad + m) * Npad + pc0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(wp + i);
      v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
    }
    if (d.act == ALDM_ACT_GEGLU) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(wp + d.bn / 2 + i);
        g[i] += t.x; g[i + 1] += t.y; g[i + 2] += t.z; g[i + 3] += t.w;
      }
    }
  }
  epi_activate(d, r, pc0, v, g);
  epi_finish(d, r, n0, 32, v, n_out);
}

// Coalesced variant for the cases the planner actually splits (no activation, alpha = 1, no accumulate; fp32,
// planes or dual output; bias / row vector / residual): one thread per (row, 4 columns), so the partial sums,
// the residual and the outputs are all full-line float4 streams.  The row-owner kernel above took 22 us per
// launch in the step's launch list (34 launches per DDIM step) for a few MB of traffic.
__global__ void __launch_bounds__(256) splitk_reduce4_kernel(const __grid_constant__ aldm_gemm_desc d, int Mpad, int Npad) {
  pdl_wait();
  const int M = d.B * d.OH * d.OW;
  const int q_per_row = (d.N + 3) >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * q_per_row) return;
  const int m = (int)(idx / q_per_row);
  const int n = (int)(idx % q_per_row) * 4;
  const RowInfo r = decode_row(d, m, M);
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < d.splitk; ++z) {          // fixed order: deterministic
    const float4 t = *reinterpret_cast<const float4*>(d.ws + ((long long)z * Mpad + m) * Npad + n);
    x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
  }
  pdl_launch();
  if (d.bias) { const float4 t = __ldg(reinterpret_cast<const float4*>(d.bias + n)); x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w; }
  if (d.rowvec) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(d.rowvec + (long long)r.b * d.ld_rowvec + n));
    x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
  }
  if (d.res) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(d.res + r.orow * d.ld_res + n));
    x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
  }
  const long long o = r.orow * d.ldo + n;
  if (d.out_mode == ALDM_OUT_F32) *reinterpret_cast<float4*>(d.out + o) = x;
  if (d.out_mode == ALDM_OUT_PLANES || d.out_hi) {
    uint2 h, l;
    split2(x.x, x.y, h.x, l.x);
    split2(x.z, x.w, h.y, l.y);
    *reinterpret_cast<uint2*>(reinterpret_cast<aldm_plane_t*>(d.out_hi) + o) = h;
    if (d.out_lo) *reinterpret_cast<uint2*>(reinterpret_cast<aldm_plane_t*>(d.out_lo) + o) = l;
  }
}

// ------------------------------------------------------------------------------------------
// SIMT checker: obviously-correct restatement of the same descriptor on CUDA cores (fp32 FMA).
// One warp per output row, lanes over 32 packed columns.  Validation / debugging only.
// ------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const __grid_constant__ aldm_gemm_desc d, int Npad) {
  const int M = d.B * d.OH * d.OW;
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const RowInfo r = decode_row(d, m, M);
  const int Hs = d.H >> d.up, Ws = d.W >> d.up;
  const int bsrc = d.bmod > 0 ? r.b % d.bmod : r.b;
  const aldm_plane_t* ahi = reinterpret_cast<const aldm_plane_t*>(d.a_hi);
  const aldm_plane_t* alo = reinterpret_cast<const aldm_plane_t*>(d.a_lo);      // NULL: single-plane operand
  const bool geglu = d.act == ALDM_ACT_GEGLU;
  const int nchunks = geglu ? (Npad / d.bn) * (d.bn / 64) : Npad / 32;
  for (int ch = blockIdx.y; ch < nchunks; ch += gridDim.y) {
    int pc0, n0, n_out;
    if (geglu) {
      const int per_tile = d.bn / 64;
      const int tile = ch / per_tile, sub = ch % per_tile;
      pc0 = tile * d.bn + sub * 32; n0 = tile * (d.bn / 2) + sub * 32; n_out = d.N / 2;
    } else {
      pc0 = ch * 32; n0 = pc0; n_out = d.N;
    }
    const float* wv = d.w_plain + (long long)(pc0 + lane) * d.Kpad;
    const float* wg = wv + (long long)(d.bn / 2) * d.Kpad;
    float acc = 0.f, accg = 0.f;
    for (int tap = 0; tap < d.ntaps; ++tap) {
      const int ih = r.oh * d.sy + d.dy[tap], iw = r.ow * d.sx + d.dx[tap];
      if (ih < 0 || ih >= d.H || iw < 0 || iw >= d.W) continue;
      const long long off = ((long long)(bsrc * Hs + (ih >> d.up)) * Ws + (iw >> d.up)) * d.Cp;
      const float* wt = wv + tap * d.Cp;
      const float* wgt = wg + tap * d.Cp;
      for (int c = 0; c < d.Cp; ++c) {
        const float a = plane_to_f(ahi[off + c]) + (alo ? plane_to_f(alo[off + c]) : 0.f);
        acc = fmaf(a, __ldg(wt + c), acc);
        if (geglu) accg = fmaf(a, __ldg(wgt + c), accg);
      }
    }
    // gather the 32 lanes' results into every lane's registers via shuffles, lane 0.. stores
    float v[32], g[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = __shfl_sync(0xffffffffu, acc, i);
      g[i] = __shfl_sync(0xffffffffu, accg, i);
    }
    if (lane == 0) {
      epi_activate(d, r, pc0, v, g);
      epi_finish(d, r, n0, 32, v, n_out);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host launch
// ------------------------------------------------------------------------------------------
static int g_num_sms = 0;

template <int BN, int EPI, int AP>
static int launch_tc3_ap(const aldm_gemm_desc& d, int M, cudaStream_t st) {
  using C = Tc3Cfg<BN, AP>;
  static bool configured = false;
  if (!configured) {
    ALDM_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc3_kernel<BN, EPI, AP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    ALDM_CHECK_CUDA(cudaGetDevice(&dev));
    ALDM_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles_m = cdiv(M, C::BM), tiles_n = cdiv(d.N, BN);
  const long long total = (long long)tiles_m * tiles_n * d.splitk;
  const int grid = (int)(total < g_num_sms ? total : g_num_sms);
  Tc3Divs fd;
  fd.ow = make_fastdiv(d.OW); fd.oh = make_fastdiv(d.OH); fd.cp = make_fastdiv(d.Cp); fd.tn = make_fastdiv(tiles_n);
  fd.bmod = make_fastdiv(d.bmod > 0 ? d.bmod : 1);
  fd.plain = d.ntaps == 1 && d.dy[0] == 0 && d.dx[0] == 0 && d.sy == 1 && d.sx == 1 && d.up == 0 && d.bmod <= 0 &&
             d.OH == d.H && d.OW == d.W;
  ALDM_CHECK_CUDA(launch_pdl(gemm_tc3_kernel<BN, EPI, AP>, dim3(grid), dim3(448), C::SMEM_BYTES, st, d, tiles_m, tiles_n, fd));
  ALDM_CHECK_CUDA(cudaGetLastError());
  if (d.splitk > 1) {
    const int Mpad = tiles_m * C::BM, Npad = tiles_n * BN;
    const int chunks = (d.act == ALDM_ACT_GEGLU) ? (Npad / BN) * (BN / 64) : Npad / 32;
    const long long tot = (long long)M * chunks;
    const bool fast = d.act == ALDM_ACT_NONE && d.alpha == 1.0f && !d.accumulate && d.N % 4 == 0 && d.ldo % 4 == 0 &&
                      (d.out_mode == ALDM_OUT_F32 || d.out_mode == ALDM_OUT_PLANES) && (!d.res || (d.ld_res % 4 == 0 && aligned16(d.res))) &&
                      (!d.bias || aligned16(d.bias)) && (!d.rowvec || (d.ld_rowvec % 4 == 0 && aligned16(d.rowvec))) &&
                      (d.out_mode != ALDM_OUT_F32 || aligned16(d.out));
    // plain launches (full serialisation): early-scheduled reduction blocks only disturbed the GEMM's last epilogue
    if (fast) {
      const long long q = (long long)M * ((d.N + 3) / 4);
      splitk_reduce4_kernel<<<(unsigned)((q + 255) / 256), 256, 0, st>>>(d, Mpad, Npad);
    } else {
      splitk_epilogue_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(d, Mpad, Npad);
    }
    ALDM_CHECK_CUDA(cudaGetLastError());
  }
  return ALDM_OK;
}

template <int BN, int EPI>
static int launch_tc3_epi(const aldm_gemm_desc& d, int M, cudaStream_t st) {
  return d.a_lo ? launch_tc3_ap<BN, EPI, 2>(d, M, st) : launch_tc3_ap<BN, EPI, 1>(d, M, st);
}

template <int BN>
static int launch_tc2(const aldm_gemm_desc& d, int M, cudaStream_t st) {
  // pick the specialised epilogue: the planner's dominant cases take the compact bodies
  const bool geglu = d.act == ALDM_ACT_GEGLU;
  const int n_out = geglu ? d.N / 2 : d.N;
  const bool co = d.splitk == 1 && n_out % 4 == 0 && d.ldo % 4 == 0 && (!d.res || d.ld_res % 4 == 0) &&
                  (d.out_mode == ALDM_OUT_F32 || d.out_mode == ALDM_OUT_PLANES || d.out_mode == ALDM_OUT_QKV);
  if (co && geglu && !d.res && d.out_mode != ALDM_OUT_QKV && BN >= 64) return launch_tc3_epi<BN, EPI_GEGLU>(d, M, st);
  const long long out_rows = (long long)d.B * d.OHF * d.OWF;
  const int ld_max = d.ldo > d.ld_res ? d.ldo : d.ld_res;
  static const bool compact_on = [] { const char* e = getenv("ALDM_EPI_COMPACT"); return !(e && e[0] == '0'); }();   // A/B switch
  const bool plain = compact_on && co && d.act == ALDM_ACT_NONE && d.alpha == 1.0f && !d.accumulate && !d.rowvec &&
                     out_rows * ld_max < (1ll << 31) &&
                     (!d.bias || aligned16(d.bias)) && (!d.res || aligned16(d.res));
  if (plain && d.out_mode == ALDM_OUT_F32 && aligned16(d.out) && (!d.out_hi || d.ldo % 4 == 0))
    return launch_tc3_epi<BN, EPI_F32N>(d, M, st);
  if (plain && d.out_mode == ALDM_OUT_PLANES) return launch_tc3_epi<BN, EPI_PLN>(d, M, st);
  if (co && d.act == ALDM_ACT_NONE) return launch_tc3_epi<BN, EPI_FAST>(d, M, st);
  return launch_tc3_epi<BN, EPI_GENERIC>(d, M, st);
}

int gemm_num_launches(const aldm_gemm_desc& d) { return ((d.impl & 0xff) != ALDM_GEMM_SIMT && d.splitk > 1) ? 2 : 1; }

int gemm_launch(const aldm_gemm_desc& d, cudaStream_t st) {
  const long long Mll = (long long)d.B * d.OH * d.OW;
  ALDM_REQUIRE(Mll > 0 && Mll < (1ll << 31), ALDM_E_SHAPE, "gemm: bad M=%lld", Mll);
  const int M = (int)Mll;
  ALDM_REQUIRE(d.bn == 32 || d.bn == 64 || d.bn == 128, ALDM_E_UNSUPPORTED, "gemm: bn=%d unsupported", d.bn);
  ALDM_REQUIRE(d.Cp % 8 == 0 && d.Cp > 0, ALDM_E_SHAPE, "gemm: Cp=%d must be a positive multiple of 8", d.Cp);
  ALDM_REQUIRE(d.ntaps >= 1 && d.ntaps <= ALDM_MAX_TAPS, ALDM_E_SHAPE, "gemm: ntaps=%d", d.ntaps);
  ALDM_REQUIRE(d.K == d.ntaps * d.Cp, ALDM_E_SHAPE, "gemm: K=%d != ntaps*Cp=%d", d.K, d.ntaps * d.Cp);
  ALDM_REQUIRE(d.Kpad % 64 == 0 && d.Kpad >= d.K, ALDM_E_SHAPE, "gemm: Kpad=%d (K=%d)", d.Kpad, d.K);
  ALDM_REQUIRE(d.N >= 1, ALDM_E_SHAPE, "gemm: N=%d", d.N);
  ALDM_REQUIRE(d.a_hi, ALDM_E_ARG, "gemm: null A plane");      // a_lo == NULL: single-plane activations
  ALDM_REQUIRE(aligned16(d.a_hi) && aligned16(d.a_lo), ALDM_E_ALIGN, "gemm: A planes not 16B aligned");
  ALDM_REQUIRE(d.up == 0 || d.up == 1, ALDM_E_ARG, "gemm: up=%d", d.up);
  ALDM_REQUIRE(d.splitk >= 1 && d.splitk <= d.Kpad / 64, ALDM_E_ARG, "gemm: splitk=%d", d.splitk);
  ALDM_REQUIRE(d.splitk == 1 || d.ws, ALDM_E_ARG, "gemm: split-K needs a workspace");
  if (d.act == ALDM_ACT_GEGLU) {
    ALDM_REQUIRE(d.bn >= 64 && d.N % d.bn == 0, ALDM_E_SHAPE, "gemm: GEGLU needs N %% bn == 0 and bn >= 64");
    ALDM_REQUIRE(!d.rowvec, ALDM_E_UNSUPPORTED, "gemm: GEGLU with rowvec");
  }
  if (d.out_mode == ALDM_OUT_QKV) {
    ALDM_REQUIRE(d.out2_hi && d.n_split > 0 && d.n_split < d.N && d.n_split % d.bn == 0 && d.n_split % 32 == 0,
                 ALDM_E_ARG, "gemm: bad QKV split (n_split=%d, N=%d, bn=%d)", d.n_split, d.N, d.bn);
    ALDM_REQUIRE(d.tok_per_batch > 0 && d.ld_t >= d.tok_per_batch && d.act != ALDM_ACT_GEGLU, ALDM_E_ARG,
                 "gemm: bad QKV token layout");
  }
  if (d.out_mode == ALDM_OUT_F32 || d.out_mode == ALDM_OUT_NCHW) {
    ALDM_REQUIRE(d.out, ALDM_E_ARG, "gemm: null out");
  } else {
    ALDM_REQUIRE(d.out_hi, ALDM_E_ARG, "gemm: null out plane");      // out_lo == NULL: single-plane output
    ALDM_REQUIRE(!d.accumulate, ALDM_E_UNSUPPORTED, "gemm: accumulate into planes");
  }
  if ((d.impl & 0xff) == ALDM_GEMM_SIMT) {
    ALDM_REQUIRE(d.w_plain, ALDM_E_ARG, "gemm: SIMT path needs w_plain");
    const int Npad = cdiv(d.N, d.bn) * d.bn;
    const bool geglu = d.act == ALDM_ACT_GEGLU;
    const int nchunks = geglu ? (Npad / d.bn) * (d.bn / 64) : Npad / 32;
    dim3 grid(cdiv(M, 4), nchunks < 64 ? nchunks : 64);
    gemm_simt_kernel<<<grid, 128, 0, st>>>(d, Npad);
    ALDM_CHECK_CUDA(cudaGetLastError());
    return ALDM_OK;
  }
  ALDM_REQUIRE(d.w_packed && aligned16(d.w_packed), ALDM_E_ARG, "gemm: w_packed null/unaligned");
  // ALDM_GEMM_TC_V1 (the round-1 one-tile-per-CTA kernel) is retired: the value selects the persistent kernel
  switch (d.bn) {
    case 128: return launch_tc2<128>(d, M, st);
    case 64: return launch_tc2<64>(d, M, st);
    default: return launch_tc2<32>(d, M, st);
  }
}

}  // namespace aldm

extern "C" int aldm_debug_timeline(long long* host_out, int32_t n) {
  using namespace aldm;
  ALDM_REQUIRE(host_out && n > 0 && n <= 4 * 256 * 2, ALDM_E_ARG, "debug_timeline: bad arguments");
  ALDM_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * n));
  void* sym = nullptr;        // cleared after every read so that stamps of different cases never mix
  ALDM_CHECK_CUDA(cudaGetSymbolAddress(&sym, g_timeline));
  ALDM_CHECK_CUDA(cudaMemset(sym, 0, sizeof(g_timeline)));
  return ALDM_OK;
}

extern "C" int aldm_gemm(const aldm_gemm_desc* d, void* stream) {
  if (!d) { aldm::set_error("aldm_gemm: null desc"); return ALDM_E_ARG; }
  return aldm::gemm_launch(*d, reinterpret_cast<cudaStream_t>(stream));
}
