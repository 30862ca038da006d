// Input-pipeline natives (reference dataset.py:12-91): the tf.data stages that touch bytes.
//   gs_crc32c            TFRecord framing checksum (tf.data.TFRecordDataset, dataset.py:46-48)
//   gs_wav_decode_pcm16  audio_ops.decode_wav(desired_channels=1, desired_samples=N) up to the int16 samples
//                        (dataset.py:28-36); the /32768 scaling to float happens on the GPU
//   gs_wav_read_batch    tf.read_file + decode_wav for a whole batch on a thread pool
//                        (dataset.map(num_parallel_calls=os.cpu_count()), dataset.py:60-63), straight into a
//                        caller-owned (pinned) int16 batch buffer
//   gs_pcm16_to_float    int16 -> float32 * (1 / 32768) on the device (decode_wav's "-32768..32767 -> -1.0..1.0")
// The first three are HOST functions (host pointers); only gs_pcm16_to_float takes device pointers.
#include <atomic>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>

#include "common.cuh"
#include "gansynth_b200.h"

namespace {

uint32_t g_crc_table[8][256];
std::atomic<int> g_crc_ready{0};

void crc_init() {
  if (g_crc_ready.load(std::memory_order_acquire)) return;
  static std::atomic<int> lock{0};
  int expected = 0;
  if (!lock.compare_exchange_strong(expected, 1)) {
    while (!g_crc_ready.load(std::memory_order_acquire)) std::this_thread::yield();
    return;
  }
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);  // Castagnoli, reflected
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
  g_crc_ready.store(1, std::memory_order_release);
}

uint32_t crc32c(const uint8_t* p, size_t n) {
  crc_init();
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {  // slicing-by-8
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = g_crc_table[7][lo & 0xFF] ^ g_crc_table[6][(lo >> 8) & 0xFF] ^ g_crc_table[5][(lo >> 16) & 0xFF] ^
        g_crc_table[4][lo >> 24] ^ g_crc_table[3][hi & 0xFF] ^ g_crc_table[2][(hi >> 8) & 0xFF] ^
        g_crc_table[1][(hi >> 16) & 0xFF] ^ g_crc_table[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_crc_table[0][(c ^ *p++) & 0xFF];
  return c ^ 0xFFFFFFFFu;
}

inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// error codes of the decoders (returned / written to status[])
enum { WAV_OK = 0, WAV_ERR_OPEN = -10, WAV_ERR_HEADER = -11, WAV_ERR_FORMAT = -12, WAV_ERR_NODATA = -13 };

// RIFF/WAVE, PCM 16-bit.  Channel 0 is kept (desired_channels = 1), the clip is cropped or zero-padded at the END to
// desired_samples.  Chunks other than "fmt " / "data" are skipped (word-aligned), like TF's DecodeLin16WaveAsFloatVector.
int wav_decode(const uint8_t* f, size_t n, int16_t* dst, int desired, int* rate, int* in_file) {
  if (n < 12 || memcmp(f, "RIFF", 4) != 0 || memcmp(f + 8, "WAVE", 4) != 0) return WAV_ERR_HEADER;
  size_t pos = 12;
  int channels = 0, bits = 0, fmt_tag = 0;
  bool have_fmt = false;
  while (pos + 8 <= n) {
    const uint8_t* ck = f + pos;
    size_t len = rd32(ck + 4);
    pos += 8;
    if (memcmp(ck, "fmt ", 4) == 0) {
      if (len < 16 || pos + 16 > n) return WAV_ERR_HEADER;
      fmt_tag = (int)rd16(f + pos);
      channels = (int)rd16(f + pos + 2);
      if (rate) *rate = (int)rd32(f + pos + 4);
      bits = (int)rd16(f + pos + 14);
      have_fmt = true;
    } else if (memcmp(ck, "data", 4) == 0) {
      if (!have_fmt) return WAV_ERR_HEADER;
      if (fmt_tag != 1 || bits != 16 || channels < 1) return WAV_ERR_FORMAT;
      if (len > n - pos) len = n - pos;  // truncated file: decode what is there
      const size_t frames = len / (2 * (size_t)channels);
      if (in_file) *in_file = (int)frames;
      const size_t take = frames < (size_t)desired ? frames : (size_t)desired;
      const uint8_t* s = f + pos;
      if (channels == 1) {
        memcpy(dst, s, take * 2);  // little-endian host
      } else {
        for (size_t i = 0; i < take; ++i) dst[i] = (int16_t)rd16(s + i * 2 * channels);
      }
      if (take < (size_t)desired) memset(dst + take, 0, ((size_t)desired - take) * 2);
      return WAV_OK;
    }
    pos += len + (len & 1);
  }
  return WAV_ERR_NODATA;
}

int wav_read_file(const char* path, int16_t* dst, int desired, std::vector<uint8_t>& buf) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return WAV_ERR_OPEN;
  // header + the samples that can be used is all that is needed; read in one go up to a generous bound
  fseek(fp, 0, SEEK_END);
  long size = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (size < 0) { fclose(fp); return WAV_ERR_OPEN; }
  buf.resize((size_t)size);
  size_t got = size ? fread(buf.data(), 1, (size_t)size, fp) : 0;
  fclose(fp);
  return wav_decode(buf.data(), got, dst, desired, nullptr, nullptr);
}

__global__ void pcm16_to_float_kernel(const short* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i8 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const float sc = 1.0f / 32768.0f;
  if (i8 + 8 <= n && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(src + i8));
    float4 a, b;
    a.x = (float)(short)(v.x & 0xFFFF) * sc; a.y = (float)(short)(v.x >> 16) * sc;
    a.z = (float)(short)(v.y & 0xFFFF) * sc; a.w = (float)(short)(v.y >> 16) * sc;
    b.x = (float)(short)(v.z & 0xFFFF) * sc; b.y = (float)(short)(v.z >> 16) * sc;
    b.z = (float)(short)(v.w & 0xFFFF) * sc; b.w = (float)(short)(v.w >> 16) * sc;
    reinterpret_cast<float4*>(dst + i8)[0] = a;
    reinterpret_cast<float4*>(dst + i8)[1] = b;
  } else {
    for (long long i = i8; i < n && i < i8 + 8; ++i) dst[i] = (float)src[i] * sc;
  }
}

}  // namespace

extern "C" int gs_crc32c(const void* data, long long n, unsigned int* out) {
  GS_CHECK_ARG(n >= 0 && out != nullptr && (data != nullptr || n == 0), "crc32c: bad arguments");
  *out = crc32c(static_cast<const uint8_t*>(data), (size_t)n);
  return GS_OK;
}

extern "C" int gs_wav_decode_pcm16(const void* file_bytes, long long n, short* dst, int desired_samples,
                                   int* sample_rate, int* samples_in_file) {
  GS_CHECK_ARG(file_bytes != nullptr && n >= 0 && dst != nullptr && desired_samples > 0, "wav_decode_pcm16: bad arguments");
  int rc = wav_decode(static_cast<const uint8_t*>(file_bytes), (size_t)n, dst, desired_samples, sample_rate, samples_in_file);
  if (rc != WAV_OK) {
    gs_set_error("wav_decode_pcm16: %s", rc == WAV_ERR_HEADER ? "not a RIFF/WAVE file or fmt chunk missing"
                                         : rc == WAV_ERR_FORMAT ? "only 16-bit PCM is supported (like audio_ops.decode_wav)"
                                                                : "no data chunk");
    return GS_ERR_ARG;
  }
  return GS_OK;
}

extern "C" int gs_wav_read_batch(const char* const* paths, int n, short* dst, int desired_samples, int threads,
                                 int* status) {
  GS_CHECK_ARG(n >= 0 && desired_samples > 0 && (n == 0 || (paths != nullptr && dst != nullptr)), "wav_read_batch: bad arguments");
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  std::atomic<int> next{0};
  std::atomic<int> failed{-1};
  auto work = [&]() {
    std::vector<uint8_t> buf;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      const int rc = wav_read_file(paths[i], dst + (size_t)i * desired_samples, desired_samples, buf);
      if (status) status[i] = rc;
      if (rc != WAV_OK) {
        memset(dst + (size_t)i * desired_samples, 0, (size_t)desired_samples * 2);
        int none = -1;
        failed.compare_exchange_strong(none, i);
      }
    }
  };
  if (threads == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  const int bad = failed.load();
  if (bad >= 0) {
    gs_set_error("wav_read_batch: cannot decode '%s' (status %d)", paths[bad], status ? status[bad] : 0);
    return GS_ERR_ARG;
  }
  return GS_OK;
}

extern "C" int gs_pcm16_to_float(const short* src, float* dst, long long n, void* stream) {
  GS_CHECK_ARG(n >= 0, "pcm16_to_float: bad size");
  if (n == 0) return GS_OK;
  const long long groups = (n + 7) / 8;
  pcm16_to_float_kernel<<<gs_cdiv(groups, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  GS_CHECK_LAUNCH("pcm16_to_float");
  return GS_OK;
}
