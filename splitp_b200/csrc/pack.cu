// Library basics + FASTA text / code matrix -> packed alignment.
// Reference semantics: splitp/parsers/fasta.py:48-57 (upper-case, site usable iff all chars in ACGT).
#include "common.cuh"
#include <string.h>

namespace spb {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace spb

using namespace spb;

extern "C" int spb_version(void) { return 100; }
extern "C" const char* spb_last_error(void) { return spb::g_err; }
extern "C" uint64_t spb_launch_count(void) { return __atomic_load_n(&spb::g_launches, __ATOMIC_RELAXED); }

extern "C" int spb_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (sms) *sms = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return 0;
}

extern "C" int64_t spb_plane_words(int64_t n_sites) {
  int64_t w = (n_sites + 31) / 32;
  return (w + 3) / 4 * 4;
}
extern "C" int64_t spb_sm_words(int n_taxa, int64_t n_sites) {
  int64_t w = (n_sites + 31) / 32 * 2 * (int64_t)n_taxa;
  return (w + 3) / 4 * 4 + 8;
}

namespace {

__device__ __forceinline__ uint32_t to_code(uint32_t c, int is_ascii) {
  if (!is_ascii) return c;
  uint32_t u = c & 0xDFu;
  return u == 'A' ? 0u : u == 'C' ? 1u : u == 'G' ? 2u : u == 'T' ? 3u : 255u;
}

// One thread = 32 consecutive sites (one plane word).  Keys are built in registers (static indices
// only), then concatenated into the site-major bit stream.
__global__ void __launch_bounds__(128) pack_kernel(const uint8_t* __restrict__ chars, int n, int64_t N, int64_t stride,
                                                   int is_ascii, uint32_t* __restrict__ sm, uint32_t* __restrict__ planes,
                                                   uint32_t* __restrict__ valid, int64_t Wp, int64_t Wn) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= Wp) return;
  int64_t s0 = w * 32;
  uint64_t key[32];
#pragma unroll
  for (int s = 0; s < 32; ++s) key[s] = 0;
  uint32_t inv = 0;
  const bool full = (s0 + 32 <= N);
  for (int j = 0; j < n; ++j) {
    const uint8_t* row = chars + (int64_t)j * stride + s0;
    uint32_t b[8];
    if (full && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
      uint4 v0 = __ldg(reinterpret_cast<const uint4*>(row));
      uint4 v1 = __ldg(reinterpret_cast<const uint4*>(row) + 1);
      b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w; b[4] = v1.x; b[5] = v1.y; b[6] = v1.z; b[7] = v1.w;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t x = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          int64_t s = s0 + q * 4 + r;
          uint32_t c = (s < N) ? (uint32_t)row[q * 4 + r] : (is_ascii ? (uint32_t)'N' : 255u);
          x |= c << (8 * r);
        }
        b[q] = x;
      }
    }
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int s = 0; s < 32; ++s) {
      uint32_t c = to_code((b[s >> 2] >> (8 * (s & 3))) & 0xFFu, is_ascii);
      uint32_t bad = c > 3u;
      inv |= bad << s;
      c &= 3u;
      if (bad) c = 0;
      lo |= (c & 1u) << s;
      hi |= (c >> 1) << s;
      key[s] = (key[s] << 2) | c;
    }
    if (planes) {
      planes[((int64_t)j * 2 + 0) * Wp + w] = lo;
      planes[((int64_t)j * 2 + 1) * Wp + w] = hi;
    }
  }
  uint32_t ok = ~inv;
  if (valid) valid[w] = ok;
  if (sm && w < Wn) {
    const int bits = 2 * n;
    uint32_t* out = sm + w * (int64_t)bits;
    unsigned __int128 acc = 0;
    int nb = 0;
#pragma unroll
    for (int s = 0; s < 32; ++s) {
      uint64_t k = ((ok >> s) & 1u) ? key[s] : 0ull;
      acc |= (unsigned __int128)k << nb;
      nb += bits;
      while (nb >= 32) {
        *out++ = (uint32_t)acc;
        acc >>= 32;
        nb -= 32;
      }
    }
  }
}

}  // namespace

extern "C" int spb_pack(const uint8_t* d_chars, int n_taxa, int64_t n_sites, int64_t row_stride, int is_ascii,
                        uint32_t* d_sm, uint32_t* d_planes, uint32_t* d_valid, void* stream) {
  SPB_REQUIRE(d_chars && n_taxa >= 1 && n_taxa <= SPB_MAX_TAXA && n_sites >= 0 && row_stride >= n_sites,
              "spb_pack: bad arguments (n_taxa=%d n_sites=%lld)", n_taxa, (long long)n_sites);
  if (d_sm && n_taxa > 32) {
    set_error("spb_pack: the site-major stream needs n_taxa <= 32 (got %d)", n_taxa);
    return SPB_ERR_UNSUPPORTED;
  }
  int64_t Wp = spb_plane_words(n_sites), Wn = (n_sites + 31) / 32;
  if (Wp == 0) return SPB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_sm) SPB_CUDA(cudaMemsetAsync(d_sm + Wn * 2 * n_taxa, 0, (spb_sm_words(n_taxa, n_sites) - Wn * 2 * n_taxa) * 4, st));
  int threads = 128;
  int64_t blocks = (Wp + threads - 1) / threads;
  pack_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_chars, n_taxa, n_sites, row_stride, is_ascii, d_sm, d_planes, d_valid, Wp, Wn);
  SPB_LAUNCH_CHECK();
  return SPB_OK;
}
