// STFT analysis and masked iSTFT synthesis around a shared-memory FFT (HBM-bound kernels).
//   analysis : util.py:171-201 (stft_mc padding rule) + librosa 0.5.1 stft(center=False) + audio_dataset.py:194
//              (sqrt-Hann) + audio_dataset.py:22-23 (magnitude);  outputs the reference's [Re;Im] stack (util.py:351)
//              and the (frames, F) magnitude rows the network consumes.
//   synthesis: audio_dataset.py:267-278 (mask on Re and Im) + util.py:48-169 (istft_noDiv: window*2/(N//hop),
//              overlap-add, no window-sum division) + util.py:219-223 (trim N both ends).
#include "internal.h"

#include <cstdlib>
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

// ---------------------------------------------------------------------------------------------------------------------
// Tiled kernels (the ones that run for the usual sizes).  A CTA handles FT consecutive frames of one utterance so that
// the reference's [Re;Im] stack - (2F x frames), FRAME index fastest - is read and written in runs of FT floats per
// bin instead of one float per bin, the audio / output samples are touched once, and two real frames share one complex
// FFT (z = x_a + i x_b) that a single warp runs in its own shared-memory buffer with warp-level barriers only.
// ---------------------------------------------------------------------------------------------------------------------

// In-place forward FFT of one warp's buffer `d` (N complex, bit-reversed order), twiddles tw[k] = exp(-2 pi i k / N).
__device__ __forceinline__ void fft_warp(float2* d, const float2* tw, int N, int logN, int lane) {
  for (int s = 1; s <= logN; ++s) {
    const int half = 1 << (s - 1);
    const int tstride = N >> s;
    __syncwarp();
    for (int i = lane; i < (N >> 1); i += 32) {
      const int j = i & (half - 1);
      const int base = ((i >> (s - 1)) << s) + j;
      const float2 w = tw[j * tstride];
      const float2 u = d[base];
      const float2 t = cmul(w, d[base + half]);
      d[base] = make_float2(u.x + t.x, u.y + t.y);
      d[base + half] = make_float2(u.x - t.x, u.y - t.y);
    }
  }
  __syncwarp();
}

// ---- four-step FFT of N = 32 * N2 points by ONE warp: lane n1 transforms x[n1 + 32 n2] over n2 in registers, the
// twiddles W_N^{n1 k2} are applied, the (N2 x 32) matrix is transposed through a padded shared-memory tile, lane k2
// transforms over n1 in registers and holds X[k2 + N2 k1].  No bit-reversal scatter, two warp barriers per transform.
constexpr __host__ __device__ int brev_c(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
constexpr __host__ __device__ int ilog2_c(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

// radix-2 DIF of M points held in registers; w[j] = exp(-2 pi i j / 32) for j < 16.  Output: X[k] = v[brev(k)].
template <int M>
__device__ __forceinline__ void fft_regs(float2 (&v)[M], const float2 (&w)[16]) {
#pragma unroll
  for (int len = M; len >= 2; len >>= 1) {
    const int half = len >> 1;
#pragma unroll
    for (int b = 0; b < M; b += len) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float2 u = v[b + j], x = v[b + j + half];
        v[b + j] = make_float2(u.x + x.x, u.y + x.y);
        const float2 dlt = make_float2(u.x - x.x, u.y - x.y);
        const int tj = j * (32 / len);                        // W_len^j = W_32^(j * 32 / len)
        v[b + j + half] = (tj == 0) ? dlt : ((tj == 8) ? make_float2(dlt.y, -dlt.x) : cmul(dlt, w[tj]));
      }
    }
  }
}

// One warp, N = 32 * N2 (N2 = 16 or 32).  `d` holds the input in natural order and receives the output in natural
// order; `tile` is the warp's transpose scratch of N2 * 33 float2; tw[k] = exp(-2 pi i k / N), k < N/2.
template <int N2>
__device__ __forceinline__ void fft_fourstep(float2* d, const float2* tw, const float2 (&w32)[16], int lane) {
  constexpr int N = 32 * N2, B2 = ilog2_c(N2);
  float2 v[N2];
#pragma unroll
  for (int n2 = 0; n2 < N2; ++n2) v[n2] = d[lane + 32 * n2];
  fft_regs<N2>(v, w32);
  __syncwarp();                                               // the transpose tile reuses the input buffer
#pragma unroll
  for (int k2 = 0; k2 < N2; ++k2) {
    const int m = lane * k2;                                  // < N
    float2 t = tw[m & (N / 2 - 1)];
    if (m >= N / 2) t = make_float2(-t.x, -t.y);
    d[k2 * 33 + lane] = cmul(v[brev_c(k2, B2)], t);
  }
  __syncwarp();
  float2 u[32];
  if (lane < N2) {
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) u[n1] = d[lane * 33 + n1];
    fft_regs<32>(u, w32);
  }
  __syncwarp();
  if (lane < N2) {
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) d[lane + N2 * k1] = u[brev_c(k1, 5)];
  }
  __syncwarp();
}

struct StftTile { int FT, pairs_a, pairs_s, halo, N2, WB; size_t smem_a, smem_s; bool ok; };

// shared-memory plan of the tiled kernels for (N, hop); ok = false -> the one-frame-per-CTA kernels are used
static StftTile stft_tile(int N, int hop) {
  StftTile t{};
  const int F = N / 2 + 1;
  t.FT = N <= 1024 ? 16 : 8;
  if (getenv("DRNMF_STFT_FT")) t.FT = atoi(getenv("DRNMF_STFT_FT"));
  t.halo = N / hop - 1;                                   // earlier frames that overlap a tile's output samples
  t.pairs_a = t.FT / 2;
  t.pairs_s = (t.FT + t.halo + 1) / 2;
  t.N2 = (N == 512 || N == 1024) ? N / 32 : 0;            // register four-step FFT; 0 = shared-memory radix-2
  t.WB = (t.N2 ? 33 * t.N2 : N) + 1;                      // float2 per warp buffer (padded transpose tile; +1: the pair
                                                          // buffers start in different banks for the cross-pair store loop)
  const size_t common = sizeof(float2) * (N / 2) + sizeof(float) * N;          // twiddles + window
  // analysis: twiddles + window + audio segment + one FFT buffer per frame pair (the [Re;Im] runs are rebuilt from the
  // buffers at store time: no staging tile, which halves the shared memory per CTA -> two CTAs per SM at N = 1024)
  t.smem_a = 16 + common + sizeof(float) * ((size_t)(t.FT - 1) * hop + N) + sizeof(float2) * (size_t)t.pairs_a * t.WB;
  t.smem_s = 16 + common + sizeof(float2) * (size_t)t.pairs_s * t.WB + 2 * sizeof(float) * (size_t)F * (2 * t.pairs_s + 1);
  t.ok = N >= 64 && N <= 2048 && t.halo <= 15 && t.pairs_s <= 32 && t.smem_a <= 200 * 1024 && t.smem_s <= 200 * 1024;
  return t;
}

// analysis: blockIdx.x = tile of FT frames, blockIdx.y = utterance, one warp per frame pair (blockDim = 32 * FT/2)
template <int N2>
__global__ void k_stft_mag_tiled(const float* __restrict__ audio, const int64_t* __restrict__ offs,
                                 const int32_t* __restrict__ lens, const int64_t* __restrict__ fidx, int N, int logN,
                                 int hop, int FT, int64_t total_frames, const float* __restrict__ wing,
                                 const float2* __restrict__ twg, float* __restrict__ stack, float* __restrict__ mag) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int F = N / 2 + 1, seg_n = (FT - 1) * hop + N, pairs = FT / 2, ld = FT + 1;
  float2* tw = reinterpret_cast<float2*>(smraw);
  float* win = reinterpret_cast<float*>(tw + N / 2);
  float* seg = win + N;
  float2* buf = reinterpret_cast<float2*>(seg + ((seg_n + 1) & ~1));
  const int WB = (N2 ? 33 * N2 : N) + 1;
  (void)pairs; (void)ld;
  const int u = blockIdx.y, j0 = blockIdx.x * FT;
  const int Tu = (int)(fidx[2 * u + 1] - fidx[2 * u]);
  if (j0 >= Tu) return;
  const int nval = min(FT, Tu - j0);
  const int64_t g0 = fidx[2 * u] + j0;
  const int len = lens[u];
  const float* src = audio + offs[u];
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) tw[k] = twg[k];
  for (int n = threadIdx.x; n < N; n += blockDim.x) win[n] = wing[n];
  for (int n0 = threadIdx.x; n0 < seg_n; n0 += 8 * blockDim.x) {   // 8 loads in flight per thread (latency-bound otherwise)
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int n = n0 + k * blockDim.x;
      const int sidx = j0 * hop + n - N;                  // N zeros in front (util.py:189-190)
      v[k] = (n < seg_n && sidx >= 0 && sidx < len) ? __ldg(src + sidx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int n = n0 + k * blockDim.x;
      if (n < seg_n) seg[n] = v[k];
    }
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ja = 2 * w, jb = 2 * w + 1;
  if (ja < nval) {
    float2* d = buf + (size_t)w * WB;
    const bool hb = jb < nval;
    for (int n = lane; n < N; n += 32) {
      const float wn = win[n];
      d[N2 ? n : (int)(__brev((unsigned)n) >> (32 - logN))] = make_float2(seg[ja * hop + n] * wn, hb ? seg[jb * hop + n] * wn : 0.f);
    }
    if constexpr (N2 != 0) {
      float2 w32[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w32[j] = tw[j * N2];
      __syncwarp();
      fft_fourstep<N2>(d, tw, w32, lane);
    } else {
      fft_warp(d, tw, N, logN, lane);
    }
    if (mag) {
      for (int f = lane; f < F; f += 32) {
        const float2 zk = d[f], zm = d[(N - f) & (N - 1)];
        // X_a = (Z[k] + conj Z[N-k]) / 2,  X_b = (Z[k] - conj Z[N-k]) / (2i);  librosa 0.5.1 conjugates the FFT
        const float are = 0.5f * (zk.x + zm.x), aim = -0.5f * (zk.y - zm.y);
        const float bre = 0.5f * (zk.y + zm.y), bim = 0.5f * (zk.x - zm.x);
        mag[(size_t)(g0 + ja) * F + f] = sqrtf(are * are + aim * aim);
        if (hb) mag[(size_t)(g0 + jb) * F + f] = sqrtf(bre * bre + bim * bim);
      }
    }
  }
  __syncthreads();
  if (stack) {
    // [Re;Im] stack: frame index fastest -> every thread rebuilds (bin f, frame j) from the pair buffer Z of frame j and
    // the CTA writes runs of FT consecutive floats per bin (the spectra stay in the FFT buffers: no staging tile)
    for (int e = threadIdx.x; e < F * FT; e += blockDim.x) {
      const int f = e / FT, j = e - f * FT;
      if (j < nval) {
        const float2* d = buf + (size_t)(j >> 1) * WB;
        const float2 zk = d[f], zm = d[(N - f) & (N - 1)];
        float re, im;
        if (j & 1) { re = 0.5f * (zk.y + zm.y); im = 0.5f * (zk.x - zm.x); }
        else       { re = 0.5f * (zk.x + zm.x); im = -0.5f * (zk.y - zm.y); }
        stack[(size_t)f * total_frames + g0 + j] = re;
        stack[(size_t)(F + f) * total_frames + g0 + j] = im;
      }
    }
  }
}

// synthesis with the overlap-add fused: a CTA owns the output samples of FT frame hops and transforms the FT + halo
// frames that touch them (one warp per frame pair, blockDim = 32 * pairs); the sum over frames runs in ascending frame
// order like the gather kernel, so the result does not depend on the tiling.
template <int N2>
__global__ void k_istft_ola_tiled(const float* __restrict__ stack, const float* __restrict__ mask,
                                  const int64_t* __restrict__ fidx, const int64_t* __restrict__ out_offs, int N, int logN,
                                  int hop, int FT, int halo, int64_t total_frames, const float* __restrict__ wing,
                                  const float2* __restrict__ twg, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int F = N / 2 + 1, pairs = (FT + halo + 1) / 2, nfr = 2 * pairs, ld = nfr + 1;
  float2* tw = reinterpret_cast<float2*>(smraw);
  float* win = reinterpret_cast<float*>(tw + N / 2);
  float2* buf = reinterpret_cast<float2*>(win + N);
  const int WB = (N2 ? 33 * N2 : N) + 1;
  float* st_re = reinterpret_cast<float*>(buf + (size_t)pairs * WB);
  float* st_im = st_re + (size_t)F * ld;
  const int u = blockIdx.y, j0 = blockIdx.x * FT;
  const int Tu = (int)(fidx[2 * u + 1] - fidx[2 * u]);
  const int out_len = hop * (Tu - 1) - N;
  if (out_len <= 0 || j0 * hop >= hop * (Tu - 1)) return;   // no output sample in this tile
  const int i0 = j0 - halo;                                  // first frame of the window (may be negative)
  const int64_t gbase = fidx[2 * u];
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) tw[k] = twg[k];
  for (int n = threadIdx.x; n < N; n += blockDim.x) win[n] = wing[n];
  for (int e = threadIdx.x; e < F * nfr; e += blockDim.x) {   // runs of nfr consecutive frames per bin
    const int f = e / nfr, j = e - f * nfr, i = i0 + j;
    const bool v = i >= 0 && i < Tu;
    st_re[f * ld + j] = v ? stack[(size_t)f * total_frames + gbase + i] : 0.f;
    st_im[f * ld + j] = v ? stack[(size_t)(F + f) * total_frames + gbase + i] : 0.f;
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int ja = 2 * w, jb = 2 * w + 1, ia = i0 + ja, ib = i0 + jb;
    const bool va = ia >= 0 && ia < Tu, vb = ib >= 0 && ib < Tu;
    float2* d = buf + (size_t)w * WB;
    if (va || vb) {
      for (int f = lane; f < F; f += 32) {
        const float ma = (va && mask) ? mask[(size_t)(gbase + ia) * F + f] : 1.f;
        const float mb = (vb && mask) ? mask[(size_t)(gbase + ib) * F + f] : 1.f;
        // (the imaginary parts at DC and Nyquist never reach the real output of the one-frame kernel; here they
        //  would leak into the partner frame, so they are dropped explicitly)
        const bool edge = (f == 0) || (f == N / 2);
        const float are = st_re[f * ld + ja] * ma, aim = edge ? 0.f : st_im[f * ld + ja] * ma;   // zero when the frame is invalid
        const float bre = st_re[f * ld + jb] * mb, bim = edge ? 0.f : st_im[f * ld + jb] * mb;
        // Z = Xext_a + i Xext_b with the Hermitian extension  Xext[N-f] = conj-part (re, -im)  of the one-frame kernel
        d[N2 ? f : (int)(__brev((unsigned)f) >> (32 - logN))] = make_float2(are - bim, aim + bre);
        if (f > 0 && f < N / 2) d[N2 ? N - f : (int)(__brev((unsigned)(N - f)) >> (32 - logN))] = make_float2(are + bim, -aim + bre);
      }
      if constexpr (N2 != 0) {
        float2 w32[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w32[j] = tw[j * N2];
        __syncwarp();
        fft_fourstep<N2>(d, tw, w32, lane);
      } else {
        fft_warp(d, tw, N, logN, lane);
      }
    }
  }
  __syncthreads();
  // overlap-add as a gather over the window's frames (ascending), window * 2/(N/hop) / N applied on the fly
  const float scale = (2.0f / (float)(N / hop)) / (float)N;      // util.py:143 window scaling, 1/N of the ifft
  float* dst = out + out_offs[u];
  for (int q = threadIdx.x; q < FT * hop; q += blockDim.x) {
    const int p = j0 * hop + q;                                  // position in the untrimmed signal
    const int sidx = p - N;
    if (sidx < 0 || sidx >= out_len) continue;
    int i_hi = p / hop; if (i_hi > Tu - 1) i_hi = Tu - 1;
    int i_lo = (p - N + hop) / hop; if (i_lo < 0) i_lo = 0;      // smallest i with p - i*hop < N
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const int j = i - i0, n = p - i * hop;
      const float2 y = buf[(size_t)(j >> 1) * WB + n];
      acc += ((j & 1) ? y.y : y.x) * win[n] * scale;
    }
    dst[sidx] = acc;
  }
}

// The same synthesis without the [Re;Im] staging tile (N = 512 / 1024, natural-order four-step FFT): the raw spectra of
// a frame pair are parked IN the pair's FFT buffer - frame a's (re, im) of bin f at d[f], frame b's at d[N-f], the
// purely real edge bins of both frames share d[0] and d[N/2] - by a pass that reads the stack in runs of consecutive
// frames; the pair's warp then applies the masks (read along f) and folds the four values of (f, N-f) into
// Z = Xext_a + i Xext_b in place.  Half the shared memory of k_istft_ola_tiled and 8 warps per CTA: two CTAs per SM.
template <int N2>
__global__ void __launch_bounds__(256, 2)
k_istft_ola_inplace(const float* __restrict__ stack, const float* __restrict__ mask, const int64_t* __restrict__ fidx,
                    const int64_t* __restrict__ out_offs, int hop, int FT, int halo, int64_t total_frames,
                    const float* __restrict__ wing, const float2* __restrict__ twg, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int N = 32 * N2, F = N / 2 + 1, WB = 33 * N2 + 1;
  constexpr int pairs = 8, nfr = 16;
  float2* tw = reinterpret_cast<float2*>(smraw);
  float* win = reinterpret_cast<float*>(tw + N / 2);
  float2* buf = reinterpret_cast<float2*>(win + N);
  const int u = blockIdx.y, j0 = blockIdx.x * FT;
  const int Tu = (int)(fidx[2 * u + 1] - fidx[2 * u]);
  const int out_len = hop * (Tu - 1) - N;
  if (out_len <= 0 || j0 * hop >= hop * (Tu - 1)) return;   // no output sample in this tile
  const int i0 = j0 - halo;                                  // first frame of the window (may be negative)
  const int64_t gbase = fidx[2 * u];
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) tw[k] = twg[k];
  for (int n = threadIdx.x; n < N; n += blockDim.x) win[n] = wing[n];
  {
    // runs of nfr = 16 consecutive frames per bin: thread t serves frame (t & 15) of the bins (t >> 4) + 16 k.  The loads
    // of UNR bins are issued before any of them is stored (the loop is latency-bound: one round trip per batch, not per bin)
    constexpr int UNR = 8;
    const int j = threadIdx.x & (nfr - 1), i = i0 + j;
    const bool v = i >= 0 && i < Tu;
    float2* d = buf + (size_t)(j >> 1) * WB;
    const float* sre = stack + gbase + (v ? i : 0);
    const float* sim = sre + (size_t)F * total_frames;
    for (int f0 = threadIdx.x >> 4; f0 < F; f0 += 16 * UNR) {
      float re[UNR], im[UNR];
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int f = f0 + 16 * k;
        const bool ok = v && f < F;
        re[k] = ok ? __ldg(sre + (size_t)f * total_frames) : 0.f;
        im[k] = (ok && f != 0 && f != N / 2) ? __ldg(sim + (size_t)f * total_frames) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int f = f0 + 16 * k;
        if (f >= F) break;
        if (f == 0 || f == N / 2) {                             // real bins (their imaginary parts never reach the output)
          float* dd = reinterpret_cast<float*>(d + f);
          dd[j & 1] = re[k];
        } else {
          d[(j & 1) ? N - f : f] = make_float2(re[k], im[k]);
        }
      }
    }
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int ia = i0 + 2 * w, ib = ia + 1;
    const bool va = ia >= 0 && ia < Tu, vb = ib >= 0 && ib < Tu;
    float2* d = buf + (size_t)w * WB;
    if (va || vb) {
      // masks of both frames, read along f: all loads of a batch of MU bins in flight before the first use
      constexpr int MU = 6;
      const float* mpa = (va && mask) ? mask + (size_t)(gbase + ia) * F : nullptr;
      const float* mpb = (vb && mask) ? mask + (size_t)(gbase + ib) * F : nullptr;
      for (int fb = lane; fb <= N / 2; fb += 32 * MU) {
        float mav[MU], mbv[MU];
#pragma unroll
        for (int k = 0; k < MU; ++k) {
          const int f = fb + 32 * k;
          mav[k] = (mpa && f <= N / 2) ? __ldg(mpa + f) : 1.f;
          mbv[k] = (mpb && f <= N / 2) ? __ldg(mpb + f) : 1.f;
        }
#pragma unroll
        for (int k = 0; k < MU; ++k) {
        const int f = fb + 32 * k;
        if (f > N / 2) break;
        const float ma = mav[k], mb = mbv[k];
        if (f == 0 || f == N / 2) {
          const float2 r = d[f];
          d[f] = make_float2(r.x * ma, r.y * mb);             // Z = are + i bre
        } else {
          const float2 ra = d[f], rb = d[N - f];
          const float are = ra.x * ma, aim = ra.y * ma, bre = rb.x * mb, bim = rb.y * mb;
          // Z = Xext_a + i Xext_b with the Hermitian extension Xext[N-f] = (re, -im)
          d[f] = make_float2(are - bim, aim + bre);
          d[N - f] = make_float2(are + bim, -aim + bre);
        }
        }
      }
      float2 w32[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w32[j] = tw[j * N2];
      __syncwarp();
      fft_fourstep<N2>(d, tw, w32, lane);
    }
  }
  __syncthreads();
  // overlap-add as a gather over the window's frames (ascending), window * 2/(N/hop) / N applied on the fly
  const float scale = (2.0f / (float)(N / hop)) / (float)N;      // util.py:143 window scaling, 1/N of the ifft
  float* dst = out + out_offs[u];
  for (int q = threadIdx.x; q < FT * hop; q += blockDim.x) {
    const int p = j0 * hop + q;                                  // position in the untrimmed signal
    const int sidx = p - N;
    if (sidx < 0 || sidx >= out_len) continue;
    int i_hi = p / hop; if (i_hi > Tu - 1) i_hi = Tu - 1;
    int i_lo = (p - N + hop) / hop; if (i_lo < 0) i_lo = 0;      // smallest i with p - i*hop < N
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const int j = i - i0, n = p - i * hop;
      const float2 y = buf[(size_t)(j >> 1) * WB + n];
      acc += ((j & 1) ? y.y : y.x) * win[n] * scale;
    }
    dst[sidx] = acc;
  }
  (void)pairs;
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
  const StftTile tl = stft_tile(N, hop);
  if (tl.ok && !getenv("DRNMF_STFT_SIMPLE")) {
    auto kern = tl.N2 == 32 ? k_stft_mag_tiled<32> : (tl.N2 == 16 ? k_stft_mag_tiled<16> : k_stft_mag_tiled<0>);
    DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem_a));
    kern<<<dim3((max_frames + tl.FT - 1) / tl.FT, n_utt), 32 * tl.pairs_a, tl.smem_a, st>>>(
        audio, offs, lens, fidx, N, ilog2(N), hop, tl.FT, total_frames, t.win, t.tw, stack, mag);
  } else {
    const int threads = N / 2 > 1024 ? 1024 : N / 2;
    const size_t smem = sizeof(float2) * (N + N / 2);
    k_stft_mag<<<dim3(max_frames, n_utt), threads, smem, st>>>(audio, offs, lens, fidx, N, ilog2(N), hop,
                                                               total_frames, t.win, t.tw, stack, mag);
  }
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
  const StftTile tl = stft_tile(N, hop);
  if (tl.ok && tl.N2 && tl.halo <= 8 && !getenv("DRNMF_STFT_SIMPLE") && !getenv("DRNMF_ISTFT_STAGED")) {
    // in-place variant: 8 frame pairs per CTA = FT output hops + halo frames
    const int FT = 16 - tl.halo;
    const size_t smem = 16 + sizeof(float2) * (N / 2) + sizeof(float) * N + sizeof(float2) * (size_t)8 * tl.WB;
    auto kern = tl.N2 == 32 ? k_istft_ola_inplace<32> : k_istft_ola_inplace<16>;
    DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((max_frames + FT - 1) / FT, n_utt), 256, smem, st>>>(stack, mask, fidx, out_offs, hop, FT, tl.halo, total_frames,
                                                                    t.win, t.tw, out_audio);
    count_launch();
  } else if (tl.ok && !getenv("DRNMF_STFT_SIMPLE")) {
    auto kern = tl.N2 == 32 ? k_istft_ola_tiled<32> : (tl.N2 == 16 ? k_istft_ola_tiled<16> : k_istft_ola_tiled<0>);
    DRNMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem_s));
    kern<<<dim3((max_frames + tl.FT - 1) / tl.FT, n_utt), 32 * tl.pairs_s, tl.smem_s, st>>>(
        stack, mask, fidx, out_offs, N, ilog2(N), hop, tl.FT, tl.halo, total_frames, t.win, t.tw, out_audio);
    count_launch();
  } else {
    const int threads = N / 2 > 1024 ? 1024 : N / 2;
    const size_t smem = sizeof(float2) * (N + N / 2);
    k_istft_frames<<<dim3(max_frames, n_utt), threads, smem, st>>>(stack, mask, fidx, N, ilog2(N), hop, total_frames,
                                                                   t.win, t.tw, frames_tmp);
    k_ola<<<dim3(64, n_utt), 256, 0, st>>>(frames_tmp, fidx, out_offs, n_utt, N, hop, out_audio);
    count_launch(2);
  }
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

}  // namespace drnmf
