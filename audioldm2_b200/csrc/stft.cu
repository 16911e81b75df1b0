// K9: STFT + mel filterbank + log for the sr_inpainting front end (stft.py:52-81,159-178).
// The reference evaluates the DFT as a conv1d with a dense [2*(N/2+1), 1, N] basis (2.15 GFLOP per
// 10 s clip), copies the result to the CPU and finishes there.  Here one CTA owns one frame:
// reflect-padded, Hann-windowed samples -> radix-2 FFT in shared memory -> magnitude -> mel GEMV
// (warp per mel row, shuffle reduction) -> log(max(., 1e-5)).  HBM traffic: 4 B/sample in (frames
// overlap in L2), 4 B * n_mels per frame out.
#include "common.cuh"

namespace aldm {

template <int LOG2N>
__global__ void __launch_bounds__(256) stft_mel_kernel(const float* __restrict__ wav, int T, int hop,
                                                       const float* __restrict__ mel_basis, int n_mels,
                                                       float* __restrict__ out, int frames, int out_frames) {
  constexpr int N = 1 << LOG2N;
  __shared__ float2 s[N];
  __shared__ float2 tw[N / 2];
  __shared__ float mag[N / 2 + 1];
  const int f = blockIdx.x, b = blockIdx.y;
  const float* w = wav + (long long)b * T;
  const int pad = N / 2;
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  // load in bit-reversed order (decimation in time)
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    int i = f * hop + n - pad;
    if (i < 0) i = -i;
    if (i >= T) i = 2 * (T - 1) - i;
    const float win = 0.5f - 0.5f * cospif(2.0f * (float)n / (float)N);   // periodic Hann (stft.py:41)
    const int r = (int)(__brev((unsigned)n) >> (32 - LOG2N));
    s[r] = make_float2(w[i] * win, 0.f);
  }
  __syncthreads();
#pragma unroll 1
  for (int st = 1; st <= LOG2N; ++st) {
    const int half = 1 << (st - 1);
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
      const int grp = k / half, j = k % half;
      const int i0 = grp * (half << 1) + j, i1 = i0 + half;
      const float2 t = tw[j << (LOG2N - st)];
      const float2 a = s[i0], c = s[i1];
      const float2 m = make_float2(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
      s[i0] = make_float2(a.x + m.x, a.y + m.y);
      s[i1] = make_float2(a.x - m.x, a.y - m.y);
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) mag[k] = sqrtf(s[k].x * s[k].x + s[k].y * s[k].y);
  __syncthreads();
  if (f >= out_frames) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = warp; m < n_mels; m += (blockDim.x >> 5)) {
    const float* bp = mel_basis + (long long)m * (N / 2 + 1);
    float acc = 0.f;
    for (int k = lane; k <= N / 2; k += 32) acc = fmaf(__ldg(bp + k), mag[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[((long long)b * out_frames + f) * n_mels + m] = logf(fmaxf(acc, 1e-5f));
  }
}

}  // namespace aldm

extern "C" int aldm_stft_mel(const float* wav, int32_t B, int32_t T, int32_t n_fft, int32_t hop, const float* mel_basis,
                             int32_t n_mels, float* out, int32_t out_frames, void* stream) {
  using namespace aldm;
  ALDM_REQUIRE(wav && mel_basis && out, ALDM_E_ARG, "stft_mel: null pointer");
  ALDM_REQUIRE(B > 0 && T > n_fft / 2 && hop > 0 && n_mels > 0, ALDM_E_SHAPE, "stft_mel: B=%d T=%d hop=%d", B, T, hop);
  const int frames = T / hop + 1;
  ALDM_REQUIRE(out_frames > 0 && out_frames <= frames, ALDM_E_SHAPE, "stft_mel: out_frames=%d > frames=%d", out_frames,
               frames);
  dim3 grid(out_frames, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (n_fft) {
    case 256: stft_mel_kernel<8><<<grid, 256, 0, st>>>(wav, T, hop, mel_basis, n_mels, out, frames, out_frames); break;
    case 512: stft_mel_kernel<9><<<grid, 256, 0, st>>>(wav, T, hop, mel_basis, n_mels, out, frames, out_frames); break;
    case 1024: stft_mel_kernel<10><<<grid, 256, 0, st>>>(wav, T, hop, mel_basis, n_mels, out, frames, out_frames); break;
    case 2048: stft_mel_kernel<11><<<grid, 256, 0, st>>>(wav, T, hop, mel_basis, n_mels, out, frames, out_frames); break;
    default: set_error("stft_mel: n_fft=%d unsupported (256/512/1024/2048)", n_fft); return ALDM_E_UNSUPPORTED;
  }
  ALDM_CHECK_CUDA(cudaGetLastError());
  return ALDM_OK;
}
