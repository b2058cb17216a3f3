// Log-mel front-end: the step right BEFORE the encoder path.
//
// The reference computes  log(clamp(MelSpectrogram(signal), 1e-5))  (processing/processor.py:53-63,155-158) with
// torchaudio's MelSpectrogram(sample_rate 16000, n_fft 400, win 400, hop 160, f_min 0, f_max 8000, n_mels 80,
// norm "slaney", mel_scale "slaney"): a centred, reflect-padded STFT with a periodic Hann window, |X|^2, and a slaney
// triangular filterbank.  torchaudio is a third-party dependency of the reference (requirements.txt:12, pinned 2.1.0);
// its published algorithm is restated by the log-mel checker under oracle/ and pinned against torchaudio itself.
//
// B200 mapping: both contractions run on the tcgen05 tensor cores as 3xTF32 GEMMs (gemm_tf32_tc): an fp32 operand x is
// split into a TF32-exact high part and the remainder, and  x.w ~= xh.wh + xl.wh + xh.wl  is ONE GEMM over the
// K-concatenated operands  [xh | xl | xh] . [wh | wh | wl]^T  with fp32 accumulation in TMEM - the spectrogram keeps
// fp32-grade accuracy (bf16 or single-pass TF32 would bury quiet bins under the rounding noise of loud ones).
//   1. logmel_frames_kernel   waveform -> windowed frames, split, [M, 3 n_fft]            (M = B * T frames)
//   2. gemm_tf32_tc           frames x [cos | -sin] DFT table -> re | im  [M, 2 NFP]
//   3. logmel_power_kernel    re^2 + im^2, split, [M, 3 NFP]
//   4. gemm_tf32_tc           power x slaney filterbank -> mel [M, n_mels]
//   5. logmel_finish_kernel   log(max(mel, 1e-5)), transposed to the reference's [B, n_mels, T]
#include <cmath>
#include <vector>

#include "module.h"

namespace cfm {
namespace {

__device__ __forceinline__ float tf32_hi(float v) {
  // nearest TF32 (10-bit mantissa): exactly representable, so the tensor core's truncation of the operand is a no-op
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

// one warp per frame
__global__ void __launch_bounds__(256) logmel_frames_kernel(const float* __restrict__ wave, const float* __restrict__ window,
                                                            float* __restrict__ frames, int B, int L, int T, int n_fft, int hop) {
  const int lane = threadIdx.x & 31;
  const int64_t m = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= static_cast<int64_t>(B) * T) return;
  const int b = static_cast<int>(m / T), t = static_cast<int>(m - static_cast<int64_t>(b) * T);
  const float* w = wave + static_cast<int64_t>(b) * L;
  float* row = frames + m * 3 * n_fft;
  const int start = t * hop - n_fft / 2;                      // centred frame (torch.stft center=True)
  for (int n = lane; n < n_fft; n += 32) {
    int i = start + n;
    if (i < 0) i = -i;                                        // pad_mode="reflect"
    if (i >= L) i = 2 * (L - 1) - i;
    const float s = w[i] * window[n];
    const float h = tf32_hi(s);
    row[n] = h;
    row[n_fft + n] = s - h;
    row[2 * n_fft + n] = h;
  }
}

// |X|^2 of the first n_freqs bins, split; thread = (frame, bin)
__global__ void __launch_bounds__(256) logmel_power_kernel(const float* __restrict__ spec, float* __restrict__ power, int64_t M,
                                                           int n_freqs, int NFP) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * NFP) return;
  const int64_t m = i / NFP;
  const int k = static_cast<int>(i - m * NFP);
  float pw = 0.f;
  if (k < n_freqs) {
    const float re = spec[m * 2 * NFP + k], im = spec[m * 2 * NFP + NFP + k];
    pw = fmaf(re, re, im * im);
  }
  const float h = tf32_hi(pw);
  float* row = power + m * 3 * NFP;
  row[k] = h;
  row[NFP + k] = pw - h;
  row[2 * NFP + k] = h;
}

// mel [B*T, n_mels] -> out [B, n_mels, T] = log(max(mel, floor)); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) logmel_finish_kernel(const float* __restrict__ mel, float* __restrict__ out, int T,
                                                            int n_mels, float floor_v) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    tile[r][tx] = (t < T && c < n_mels) ? mel[(static_cast<int64_t>(b) * T + t) * n_mels + c] : 1.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (c < n_mels && t < T) out[(static_cast<int64_t>(b) * n_mels + c) * T + t] = logf(fmaxf(tile[tx][r], floor_v));
  }
}

__global__ void logmel_lengths_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out, int B, int hop) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) out[i] = in[i] / hop + 1;                        // processor.py:392
}

// slaney mel scale (torchaudio.functional.functional._hz_to_mel / _mel_to_hz, mel_scale="slaney")
double hz_to_mel_slaney(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz_slaney(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
double hz_to_mel_htk(double f) { return 2595.0 * std::log10(1.0 + f / 700.0); }
double mel_to_hz_htk(double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); }

inline float tf32_hi_host(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
// row-major [rows, 3*K] = [hi | hi | lo] of a [rows, K] fp64 table (the W side of the 3xTF32 product)
void split_rows(const std::vector<double>& src, int rows, int K, std::vector<float>& dst) {
  dst.assign(static_cast<size_t>(rows) * 3 * K, 0.f);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      const float v = static_cast<float>(src[static_cast<size_t>(r) * K + k]);
      const float h = tf32_hi_host(v);
      float* row = dst.data() + static_cast<size_t>(r) * 3 * K;
      row[k] = h;
      row[K + k] = h;
      row[2 * K + k] = v - h;
    }
}

}  // namespace

struct LogMel {
  cfm_logmel_config_t cfg{};
  int n_freqs = 0, NFP = 0, num_sms = 148;
  float *window = nullptr, *dft = nullptr, *fb = nullptr;
  ~LogMel() {
    if (window) cudaFree(window);
    if (dft) cudaFree(dft);
    if (fb) cudaFree(fb);
  }
  int init(const float* window_host, const float* fb_host);
  int64_t workspace_bytes(int B, int L) const;
  int forward(const float* wave, float* out, int B, int L, void* ws, int64_t ws_bytes, cudaStream_t st) const;
};

int LogMel::init(const float* window_host, const float* fb_host) {
  const int n_fft = cfg.n_fft, win = cfg.win_length, n_mels = cfg.n_mels;
  CFM_REQUIRE(cfg.sample_rate > 0 && n_fft >= 8 && n_fft % 4 == 0 && win > 0 && win <= n_fft && cfg.hop_length > 0 && n_mels > 0 &&
                  n_mels % 4 == 0, "bad log-mel configuration (n_fft and n_mels must be multiples of 4)");
  CFM_TRY(tc_init());
  int dev = 0;
  CFM_CUDA_OK(cudaGetDevice(&dev));
  CFM_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  n_freqs = n_fft / 2 + 1;
  NFP = (n_freqs + 7) & ~7;
  const double PI = 3.14159265358979323846;
  // periodic Hann window of win_length, centred in n_fft (torch.stft pads a shorter window on both sides)
  std::vector<float> w(n_fft, 0.f);
  const int left = (n_fft - win) / 2;
  for (int n = 0; n < win; ++n) w[left + n] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * PI * n / win));
  if (window_host != nullptr)                                   // the caller's window buffer (already n_fft long)
    for (int n = 0; n < n_fft; ++n) w[n] = window_host[n];
  // DFT table as a [2 NFP, n_fft] "weight": rows [0, n_freqs) cos(2 pi k n / N), rows [NFP, NFP + n_freqs) -sin(...)
  std::vector<double> tab(static_cast<size_t>(2) * NFP * n_fft, 0.0);
  for (int k = 0; k < n_freqs; ++k)
    for (int n = 0; n < n_fft; ++n) {
      const double ang = 2.0 * PI * static_cast<double>((static_cast<int64_t>(k) * n) % n_fft) / n_fft;
      tab[static_cast<size_t>(k) * n_fft + n] = std::cos(ang);
      tab[static_cast<size_t>(NFP + k) * n_fft + n] = -std::sin(ang);
    }
  std::vector<float> tab3;
  split_rows(tab, 2 * NFP, n_fft, tab3);
  // slaney / htk triangular filterbank, torchaudio.functional.melscale_fbanks: fb[f, m], stored here as [n_mels, NFP]
  const bool slaney = cfg.mel_scale_slaney != 0;
  const double f_max = cfg.f_max > 0 ? cfg.f_max : cfg.sample_rate / 2.0;
  const double m_min = slaney ? hz_to_mel_slaney(cfg.f_min) : hz_to_mel_htk(cfg.f_min);
  const double m_max = slaney ? hz_to_mel_slaney(f_max) : hz_to_mel_htk(f_max);
  std::vector<double> f_pts(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) {
    const double m = m_min + (m_max - m_min) * i / (n_mels + 1);
    f_pts[i] = slaney ? mel_to_hz_slaney(m) : mel_to_hz_htk(m);
  }
  std::vector<double> fbm(static_cast<size_t>(n_mels) * NFP, 0.0);
  for (int m = 0; m < n_mels; ++m) {
    const double enorm = cfg.norm_slaney ? 2.0 / (f_pts[m + 2] - f_pts[m]) : 1.0;
    for (int k = 0; k < n_freqs; ++k) {
      const double f = (cfg.sample_rate / 2.0) * k / (n_freqs - 1);
      const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
      fbm[static_cast<size_t>(m) * NFP + k] = std::fmax(0.0, std::fmin(down, up)) * enorm;
    }
  }
  if (fb_host != nullptr)                                       // the caller's filterbank buffer, [n_freqs, n_mels]
    for (int m = 0; m < n_mels; ++m)
      for (int k = 0; k < n_freqs; ++k) fbm[static_cast<size_t>(m) * NFP + k] = fb_host[static_cast<size_t>(k) * n_mels + m];
  std::vector<float> fb3;
  split_rows(fbm, n_mels, NFP, fb3);
  CFM_CUDA_OK(cudaMalloc(&window, w.size() * 4));
  CFM_CUDA_OK(cudaMalloc(&dft, tab3.size() * 4));
  CFM_CUDA_OK(cudaMalloc(&fb, fb3.size() * 4));
  CFM_CUDA_OK(cudaMemcpy(window, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CFM_CUDA_OK(cudaMemcpy(dft, tab3.data(), tab3.size() * 4, cudaMemcpyHostToDevice));
  CFM_CUDA_OK(cudaMemcpy(fb, fb3.data(), fb3.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

int64_t LogMel::workspace_bytes(int B, int L) const {
  const int64_t M = static_cast<int64_t>(B) * (L / cfg.hop_length + 1);
  // frames [M, 3 n_fft] (re-used for the split power [M, 3 NFP] and mel [M, n_mels]) + spectrum [M, 2 NFP]
  return M * 3 * cfg.n_fft * 4 + M * 2 * NFP * 4 + 1024;
}

int LogMel::forward(const float* wave, float* out, int B, int L, void* ws, int64_t ws_bytes, cudaStream_t st) const {
  CFM_REQUIRE(wave != nullptr && out != nullptr && B > 0, "bad log-mel arguments");
  CFM_REQUIRE(L > cfg.n_fft / 2, "waveform shorter than the reflect padding (n_fft / 2 samples)");
  CFM_REQUIRE(ws != nullptr && ws_bytes >= workspace_bytes(B, L), "workspace too small; query cfm_logmel_workspace_bytes");
  const int T = L / cfg.hop_length + 1, n_fft = cfg.n_fft;
  const int64_t M = static_cast<int64_t>(B) * T;
  CFM_REQUIRE(M < (1ll << 31), "too many frames");
  float* frames = static_cast<float*>(ws);
  float* spec = frames + M * 3 * n_fft;
  float* power = frames;                                   // frames are dead after the DFT
  float* mel = frames + M * 3 * NFP;                       // 3 NFP + n_mels <= 3 n_fft floats per frame
  CFM_REQUIRE(3 * NFP + cfg.n_mels <= 3 * n_fft, "n_mels too large for the workspace layout");
  logmel_frames_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, st>>>(wave, window, frames, B, L, T, n_fft, cfg.hop_length);
  CFM_CUDA_OK(cudaGetLastError());
  CFM_TRY(gemm_tf32_tc(frames, 3 * n_fft, dft, 3 * n_fft, nullptr, spec, 2 * NFP, static_cast<int>(M), 2 * NFP, 3 * n_fft, num_sms, st));
  logmel_power_kernel<<<static_cast<unsigned>((M * NFP + 255) / 256), 256, 0, st>>>(spec, power, M, n_freqs, NFP);
  CFM_CUDA_OK(cudaGetLastError());
  CFM_TRY(gemm_tf32_tc(power, 3 * NFP, fb, 3 * NFP, nullptr, mel, cfg.n_mels, static_cast<int>(M), cfg.n_mels, 3 * NFP, num_sms, st));
  dim3 grid(static_cast<unsigned>((T + 31) / 32), static_cast<unsigned>((cfg.n_mels + 31) / 32), static_cast<unsigned>(B));
  logmel_finish_kernel<<<grid, 256, 0, st>>>(mel, out, T, cfg.n_mels, cfg.log_floor);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm

struct cfm_logmel : public cfm::LogMel {};

using namespace cfm;

extern "C" {

int cfm_logmel_create(const cfm_logmel_config_t* cfg, const float* window_host, const float* fb_host, cfm_logmel_t* out) {
  CFM_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  cfm_logmel* h = new (std::nothrow) cfm_logmel();
  CFM_REQUIRE(h != nullptr, "out of host memory");
  h->cfg = *cfg;
  const int s = h->init(window_host, fb_host);
  if (s != 0) {
    delete h;
    return s;
  }
  *out = h;
  return 0;
}
int cfm_logmel_destroy(cfm_logmel_t h) {
  delete h;
  return 0;
}
int64_t cfm_logmel_workspace_bytes(cfm_logmel_t h, int32_t batch, int32_t samples) {
  if (h == nullptr || batch <= 0 || samples <= 0) { set_error("bad log-mel workspace query"); return -1; }
  return h->workspace_bytes(batch, samples);
}
int cfm_logmel_forward(cfm_logmel_t h, const float* wave, float* mel, int32_t batch, int32_t samples, void* workspace,
                       int64_t workspace_bytes, void* stream) {
  CFM_REQUIRE(h != nullptr, "null log-mel handle");
  return h->forward(wave, mel, batch, samples, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
int cfm_logmel_lengths(const int64_t* samples, int64_t* frames, int32_t batch, int32_t hop_length, void* stream) {
  CFM_REQUIRE(samples != nullptr && frames != nullptr && batch > 0 && hop_length > 0, "bad lengths arguments");
  logmel_lengths_kernel<<<(batch + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(samples, frames, batch, hop_length);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
