// K5 (first version): fused softmax(scale * Q K^T + mask) V, head_dim 32, flash-style online softmax.
// One thread owns one query row (q and the output accumulator live in registers); K/V tiles of
// 64 keys are staged in shared memory and broadcast-read.  No N x N score matrix touches HBM
// (the reference materialises it: attention.py:354-366).  fp32 FMA on CUDA cores -- exact w.r.t.
// the oracle; the tcgen05 version replaces the two inner products.
//
// Also: row softmax for the VAE AttnBlock (model.py:216-217) whose scores come from the GEMM.
#include <float.h>

#include "common.cuh"

namespace aldm {

static constexpr int ATT_D = 32;
static constexpr int ATT_KT = 64;   // keys per shared-memory tile

__global__ void __launch_bounds__(128) attention_kernel(const __grid_constant__ aldm_attn_desc d) {
  __shared__ __align__(16) float sK[ATT_KT][ATT_D];
  __shared__ __align__(16) float sV[ATT_KT][ATT_D];
  __shared__ float sM[ATT_KT];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  const bool qvalid = qi < d.Nq;
  const int bkv = d.kv_bmod > 0 ? b % d.kv_bmod : b;
  float q[ATT_D], o[ATT_D];
  {
    const float* qp = d.q + ((long long)b * d.Nq + (qvalid ? qi : 0)) * d.ldq + h * ATT_D;
#pragma unroll
    for (int i = 0; i < ATT_D; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qp + i);
      q[i] = t.x * d.scale; q[i + 1] = t.y * d.scale; q[i + 2] = t.z * d.scale; q[i + 3] = t.w * d.scale;
    }
  }
#pragma unroll
  for (int i = 0; i < ATT_D; ++i) o[i] = 0.f;
  float mrun = -INFINITY, lrun = 0.f;

  for (int k0 = 0; k0 < d.Nk; k0 += ATT_KT) {
    const int kn = min(ATT_KT, d.Nk - k0);
    __syncthreads();
    // cooperative load: 64 keys x 32 floats = 512 float4 for K and for V
    for (int idx = threadIdx.x; idx < ATT_KT * (ATT_D / 4); idx += blockDim.x) {
      const int kk = idx / (ATT_D / 4), c4 = (idx % (ATT_D / 4)) * 4;
      float4 tk = make_float4(0.f, 0.f, 0.f, 0.f), tv = tk;
      if (kk < kn) {
        const long long rowk = (long long)bkv * d.Nk + k0 + kk;
        tk = *reinterpret_cast<const float4*>(d.k + rowk * d.ldk + h * ATT_D + c4);
        tv = *reinterpret_cast<const float4*>(d.v + rowk * d.ldv + h * ATT_D + c4);
      }
      *reinterpret_cast<float4*>(&sK[kk][c4]) = tk;
      *reinterpret_cast<float4*>(&sV[kk][c4]) = tv;
    }
    for (int kk = threadIdx.x; kk < ATT_KT; kk += blockDim.x)
      sM[kk] = (d.mask && kk < kn) ? d.mask[(long long)bkv * d.Nk + k0 + kk] : 1.0f;
    __syncthreads();

    for (int c0 = 0; c0 < kn; c0 += 8) {
      float s[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int kk = c0 + jj;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < ATT_D; i += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&sK[kk < kn ? kk : 0][i]);
          acc = fmaf(q[i], t.x, acc); acc = fmaf(q[i + 1], t.y, acc);
          acc = fmaf(q[i + 2], t.z, acc); acc = fmaf(q[i + 3], t.w, acc);
        }
        if (kk >= kn) acc = -INFINITY;                     // beyond the key range: excluded
        else if (sM[kk] != 1.0f) acc = -FLT_MAX;           // masked_fill(-finfo.max), attention.py:356-360
        s[jj] = acc;
        cmax = fmaxf(cmax, acc);
      }
      const float mnew = fmaxf(mrun, cmax);
      const float corr = (mrun == -INFINITY) ? 0.f : expf(mrun - mnew);
      lrun *= corr;
#pragma unroll
      for (int i = 0; i < ATT_D; ++i) o[i] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int kk = c0 + jj;
        const float p = (s[jj] == -INFINITY) ? 0.f : expf(s[jj] - mnew);
        lrun += p;
#pragma unroll
        for (int i = 0; i < ATT_D; i += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&sV[kk < kn ? kk : 0][i]);
          o[i] = fmaf(p, t.x, o[i]); o[i + 1] = fmaf(p, t.y, o[i + 1]);
          o[i + 2] = fmaf(p, t.z, o[i + 2]); o[i + 3] = fmaf(p, t.w, o[i + 3]);
        }
      }
      mrun = mnew;
    }
  }
  if (!qvalid) return;
  const float inv = 1.0f / lrun;
#pragma unroll
  for (int i = 0; i < ATT_D; ++i) o[i] *= inv;
  const long long orow = (long long)b * d.Nq + qi;
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

int attention_launch(const aldm_attn_desc& d, cudaStream_t st) {
  ALDM_REQUIRE(d.q && d.k && d.v && d.out_hi && d.out_lo, ALDM_E_ARG, "attention: null pointer");
  ALDM_REQUIRE(d.B > 0 && d.heads > 0 && d.Nq > 0 && d.Nk > 0, ALDM_E_SHAPE, "attention: B=%d heads=%d Nq=%d Nk=%d", d.B,
               d.heads, d.Nq, d.Nk);
  ALDM_REQUIRE(d.ldq % 4 == 0 && d.ldk % 4 == 0 && d.ldv % 4 == 0 && d.ldo % 8 == 0, ALDM_E_ALIGN,
               "attention: leading dims must be multiples of 4 (ldo of 8)");
  ALDM_REQUIRE(aligned16(d.q) && aligned16(d.k) && aligned16(d.v) && aligned16(d.out_hi) && aligned16(d.out_lo),
               ALDM_E_ALIGN, "attention: pointers must be 16B aligned");
  ALDM_REQUIRE(d.heads <= 65535 && d.B <= 65535, ALDM_E_SHAPE, "attention: grid too large");
  const int threads = d.Nq >= 128 ? 128 : ((d.Nq + 31) / 32) * 32;
  dim3 grid(cdiv(d.Nq, threads), d.heads, d.B);
  attention_kernel<<<grid, threads, 0, st>>>(d);
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
