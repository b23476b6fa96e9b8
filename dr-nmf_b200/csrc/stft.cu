// STFT analysis and masked iSTFT synthesis around a shared-memory FFT (HBM-bound kernels).
//   analysis : util.py:171-201 (stft_mc padding rule) + librosa 0.5.1 stft(center=False) + audio_dataset.py:194
//              (sqrt-Hann) + audio_dataset.py:22-23 (magnitude);  outputs the reference's [Re;Im] stack (util.py:351)
//              and the (frames, F) magnitude rows the network consumes.
//   synthesis: audio_dataset.py:267-278 (mask on Re and Im) + util.py:48-169 (istft_noDiv: window*2/(N//hop),
//              overlap-add, no window-sum division) + util.py:219-223 (trim N both ends).
#include "internal.h"

#include <map>

namespace drnmf {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place forward FFT of `d` (N complex, already in bit-reversed order) with twiddle table tw[k] = exp(-2 pi i k/N).
__device__ __forceinline__ void fft_smem(float2* d, const float2* tw, int N, int logN) {
  for (int s = 1; s <= logN; ++s) {
    const int half = 1 << (s - 1);
    const int tstride = N >> s;
    __syncthreads();
    for (int i = threadIdx.x; i < (N >> 1); i += blockDim.x) {
      const int j = i & (half - 1);
      const int base = ((i >> (s - 1)) << s) + j;
      const float2 w = tw[j * tstride];
      const float2 u = d[base];
      const float2 t = cmul(w, d[base + half]);
      d[base] = make_float2(u.x + t.x, u.y + t.y);
      d[base + half] = make_float2(u.x - t.x, u.y - t.y);
    }
  }
  __syncthreads();
}

// window table: win[n] = sqrt(float32(hann_periodic(n)))  (audio_dataset.py:194), twiddles tw[k] = exp(-2 pi i k / N)
__global__ void k_fft_tables(int N, float* win, float2* tw) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) {
    float hann = (float)(0.5 - 0.5 * cospi(2.0 * (double)n / (double)N));
    win[n] = sqrtf(hann);
  }
  if (n < N / 2) {
    double s, c;
    sincospi(-2.0 * (double)n / (double)N, &s, &c);
    tw[n] = make_float2((float)c, (float)s);
  }
}

struct FftTables { float* win; float2* tw; };
static std::map<std::pair<int, int>, FftTables> g_tables;   // (device, N) -> tables (library-owned, a few KB each)

static int get_tables(int N, cudaStream_t st, FftTables* out) {
  int dev = 0;
  DRNMF_CUDA(cudaGetDevice(&dev));
  auto key = std::make_pair(dev, N);
  auto it = g_tables.find(key);
  if (it == g_tables.end()) {
    FftTables t;
    DRNMF_CUDA(cudaMalloc(&t.win, sizeof(float) * N));
    DRNMF_CUDA(cudaMalloc(&t.tw, sizeof(float2) * (N / 2)));
    k_fft_tables<<<(N + 255) / 256, 256, 0, st>>>(N, t.win, t.tw);
    count_launch();
    DRNMF_CUDA(cudaGetLastError());
    it = g_tables.emplace(key, t).first;
  }
  *out = it->second;
  return DRNMF_OK;
}

// one CTA per frame (blockIdx.x = frame within utterance blockIdx.y); smem: N float2 data + N/2 float2 twiddles
__global__ void k_stft_mag(const float* __restrict__ audio, const int64_t* __restrict__ offs,
                           const int32_t* __restrict__ lens, const int64_t* __restrict__ fidx, int N,
                           int logN, int hop, int64_t total_frames, const float* __restrict__ win,
                           const float2* __restrict__ twg, float* __restrict__ stack, float* __restrict__ mag) {
  extern __shared__ float2 sm[];
  float2* d = sm;
  float2* tw = sm + N;
  const int u = blockIdx.y;
  const int i = blockIdx.x;
  if (i >= (int)(fidx[2 * u + 1] - fidx[2 * u])) return;
  const int64_t g = fidx[2 * u] + i;
  const int len = lens[u];
  const float* src = audio + offs[u];
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) tw[k] = twg[k];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int s = i * hop + n - N;                       // N zeros in front (util.py:189-190)
    const float v = (s >= 0 && s < len) ? src[s] * win[n] : 0.f;
    d[__brev((unsigned)n) >> (32 - logN)] = make_float2(v, 0.f);
  }
  fft_smem(d, tw, N, logN);
  const int F = N / 2 + 1;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float re = d[f].x, im = -d[f].y;               // librosa 0.5.1 conjugates the FFT
    if (stack) {
      stack[(size_t)f * total_frames + g] = re;
      stack[(size_t)(F + f) * total_frames + g] = im;
    }
    if (mag) mag[(size_t)g * F + f] = sqrtf(re * re + im * im);
  }
}

// one CTA per frame: masked spectrum -> Hermitian extension -> FFT -> real/N * synthesis window -> frames_tmp
__global__ void k_istft_frames(const float* __restrict__ stack, const float* __restrict__ mask,
                               const int64_t* __restrict__ fidx, int N, int logN,
                               int hop, int64_t total_frames, const float* __restrict__ win,
                               const float2* __restrict__ twg, float* __restrict__ frames_tmp) {
  extern __shared__ float2 sm[];
  float2* d = sm;
  float2* tw = sm + N;
  const int u = blockIdx.y;
  if ((int)blockIdx.x >= (int)(fidx[2 * u + 1] - fidx[2 * u])) return;
  const int64_t g = fidx[2 * u] + blockIdx.x;
  const int F = N / 2 + 1;
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) tw[k] = twg[k];
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float mk = mask ? mask[(size_t)g * F + f] : 1.f;
    const float re = stack[(size_t)f * total_frames + g] * mk;
    const float im = stack[(size_t)(F + f) * total_frames + g] * mk;
    d[__brev((unsigned)f) >> (32 - logN)] = make_float2(re, im);
    if (f > 0 && f < N / 2) d[__brev((unsigned)(N - f)) >> (32 - logN)] = make_float2(re, -im);
  }
  fft_smem(d, tw, N, logN);
  const float scale = (2.0f / (float)(N / hop)) / (float)N;      // util.py:143 window scaling, 1/N of the ifft
  for (int n = threadIdx.x; n < N; n += blockDim.x) frames_tmp[(size_t)g * N + n] = d[n].x * win[n] * scale;
}

// overlap-add as a gather (deterministic): out[u][s] = sum_i frames_tmp[foffs[u]+i][s + N - i*hop]
__global__ void k_ola(const float* __restrict__ frames_tmp, const int64_t* __restrict__ fidx,
                      const int64_t* __restrict__ out_offs, int n_utt, int N, int hop, float* __restrict__ out) {
  const int u = blockIdx.y;
  const int T = (int)(fidx[2 * u + 1] - fidx[2 * u]);
  const int out_len = hop * (T - 1) - N;
  const float* ft = frames_tmp + (size_t)fidx[2 * u] * N;
  float* dst = out + out_offs[u];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < out_len; s += gridDim.x * blockDim.x) {
    const int p = s + N;                                  // position in the untrimmed signal
    int i_hi = p / hop; if (i_hi > T - 1) i_hi = T - 1;
    int i_lo = (p - N + hop) / hop; if (i_lo < 0) i_lo = 0;   // smallest i with p - i*hop < N
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) acc += ft[(size_t)i * N + (p - i * hop)];
    dst[s] = acc;
  }
}

static int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

int launch_stft_mag(const float* audio, const int64_t* offs, const int32_t* lens, const int64_t* fidx, int n_utt,
                    int max_frames, int N, int hop, int64_t total_frames, float* stack, float* mag, cudaStream_t st) {
  DRNMF_CHECK(N >= 32 && N <= 4096 && (N & (N - 1)) == 0, "STFT size N=%d must be a power of two in [32, 4096]", N);
  DRNMF_CHECK(hop > 0 && N % hop == 0, "hop=%d must divide N=%d", hop, N);
  if (total_frames == 0 || n_utt == 0 || max_frames == 0) return DRNMF_OK;
  FftTables t;
  int rc = get_tables(N, st, &t);
  if (rc) return rc;
  const int threads = N / 2 > 1024 ? 1024 : N / 2;
  const size_t smem = sizeof(float2) * (N + N / 2);
  k_stft_mag<<<dim3(max_frames, n_utt), threads, smem, st>>>(audio, offs, lens, fidx, N, ilog2(N), hop,
                                                             total_frames, t.win, t.tw, stack, mag);
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

int launch_mask_istft(const float* stack, const float* mask, const int64_t* fidx, const int64_t* out_offs, int n_utt,
                      int max_frames, int N, int hop, int64_t total_frames, float* frames_tmp, float* out_audio,
                      cudaStream_t st) {
  DRNMF_CHECK(N >= 32 && N <= 4096 && (N & (N - 1)) == 0, "STFT size N=%d must be a power of two in [32, 4096]", N);
  DRNMF_CHECK(hop > 0 && N % hop == 0, "hop=%d must divide N=%d", hop, N);
  if (total_frames == 0 || n_utt == 0 || max_frames == 0) return DRNMF_OK;
  FftTables t;
  int rc = get_tables(N, st, &t);
  if (rc) return rc;
  const int threads = N / 2 > 1024 ? 1024 : N / 2;
  const size_t smem = sizeof(float2) * (N + N / 2);
  k_istft_frames<<<dim3(max_frames, n_utt), threads, smem, st>>>(stack, mask, fidx, N, ilog2(N), hop, total_frames,
                                                                 t.win, t.tw, frames_tmp);
  k_ola<<<dim3(64, n_utt), 256, 0, st>>>(frames_tmp, fidx, out_offs, n_utt, N, hop, out_audio);
  count_launch(2);
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

}  // namespace drnmf
