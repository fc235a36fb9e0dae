// Marginals of a flattening, its rank-1 (independence) approximation and the rank-1 divergence
//   D(F) = sum over non-zero cells of F[x,y] * log(F[x,y] / (r[y] * c[x])),  r = column sums, c = row sums
// (reference: splitp/phylogenetics.py:331-373; banned-pattern marginals: splitp/constructions.py:94-101 as used by
// phylogenetics.py:344-361).  Two routes: from a materialised matrix (any flattening handed to the drop-in API), and
// straight from the pattern table for a split that covers all taxa, where every pattern is its own cell and nothing
// needs to be materialised.  Side sums live in a direct-indexed array (4^side doubles) or, for long sides, in an
// open-addressing table keyed by the side index.
#include "common.cuh"

namespace spb {
namespace {

constexpr int kThreads = 256;
constexpr uint64_t kEmptyKey = ~0ull;

inline int nblk(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// fixed-order tree: the result depends on the inputs only, not on scheduling
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0;
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

// ---- matrix route -------------------------------------------------------------------------------------------------
// one block per row, threads strided over the columns
__global__ void rowsum_kernel(const double* __restrict__ F, int64_t rows, int64_t cols, int64_t ld, double* __restrict__ rowsum) {
  __shared__ double sh[kThreads / 32];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const double* row = F + r * ld;
    double acc = 0.0;
    for (int64_t c = threadIdx.x; c < cols; c += kThreads) acc += row[c];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) rowsum[r] = acc;
  }
}

// block = 32 columns x 8 row lanes: coalesced 256-byte row segments
__global__ void colsum_kernel(const double* __restrict__ F, int64_t rows, int64_t cols, int64_t ld, double* __restrict__ colsum) {
  __shared__ double sh[8][33];
  const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
  double acc = 0.0;
  if (c < cols)
    for (int64_t r = threadIdx.y; r < rows; r += 8) acc += F[r * ld + c];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][threadIdx.x];
    colsum[c] = s;
  }
}

__global__ void outer_kernel(const double* __restrict__ x, int64_t nx, const double* __restrict__ y, int64_t ny,
                             double* __restrict__ out, int accumulate) {
  const int64_t total = nx * ny;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i / ny] * y[i % ny];
    out[i] = accumulate ? out[i] + v : v;
  }
}

__global__ void dense_terms_kernel(const double* __restrict__ F, int64_t rows, int64_t cols, int64_t ld,
                                   const double* __restrict__ rowsum, const double* __restrict__ colsum,
                                   double* __restrict__ partials) {
  __shared__ double sh[kThreads / 32];
  double acc = 0.0;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const double* row = F + r * ld;
    const double cr = rowsum[r];
    for (int64_t c = threadIdx.x; c < cols; c += kThreads) {
      const double v = row[c];
      if (v != 0.0) acc += v * log(v / (colsum[c] * cr));
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double sh[kThreads / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) acc += partials[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = acc;
}

// ---- pattern-table route --------------------------------------------------------------------------------------------
struct SideMap {
  double* sum;
  uint64_t* hkeys;  // nullptr: sum is indexed by the side index itself
  uint64_t mask;    // capacity - 1 of the open-addressing table
};

__device__ __forceinline__ double* slot_insert(const SideMap& m, uint64_t idx, uint32_t* overflow) {
  if (!m.hkeys) return m.sum + idx;
  uint64_t s = mix64(idx) & m.mask;
  for (uint64_t probes = 0; probes <= m.mask; ++probes) {
    uint64_t k = m.hkeys[s];
    if (k == kEmptyKey) {
      k = atomicCAS((unsigned long long*)&m.hkeys[s], (unsigned long long)kEmptyKey, (unsigned long long)idx);
      if (k == kEmptyKey) return m.sum + s;
    }
    if (k == idx) return m.sum + s;
    s = (s + 1) & m.mask;
  }
  atomicExch(overflow, 1u);
  return nullptr;
}

__device__ __forceinline__ double slot_find(const SideMap& m, uint64_t idx) {
  if (!m.hkeys) return m.sum[idx];
  uint64_t s = mix64(idx) & m.mask;
  for (uint64_t probes = 0; probes <= m.mask; ++probes) {
    const uint64_t k = m.hkeys[s];
    if (k == idx) return m.sum[s];
    if (k == kEmptyKey) break;
    s = (s + 1) & m.mask;
  }
  return 0.0;
}

__device__ __forceinline__ int digit_count(uint64_t index, int digits, int code) {
  int c = 0;
  for (int d = 0; d < digits; ++d) c += (int)((index >> (2 * d)) & 3ull) == code;
  return c;
}

__device__ __forceinline__ double raw_value(const void* vals, int kind, int64_t i) {
  return kind == SPB_VAL_U32 ? (double)((const uint32_t*)vals)[i] : ((const double*)vals)[i];
}

// Side sums of the raw values (counts are summed as doubles: integer sums below 2^53 are exact in any order, so the
// count route is deterministic).  Short sides are first accumulated in shared memory, one private copy per block.
__global__ void table_marginals_kernel(const uint64_t* __restrict__ keys, const void* __restrict__ vals, int kind, int64_t num,
                                       const __grid_constant__ SplitDev sp, int ban_row, int ban_col, SideMap rm, SideMap cm,
                                       int smem_r, int smem_c, uint32_t* overflow) {
  extern __shared__ double smem[];
  double* sr = smem;
  double* sc = smem + smem_r;
  for (int i = threadIdx.x; i < smem_r + smem_c; i += blockDim.x) smem[i] = 0.0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    const uint64_t row = side_index(key, sp.sh_a, sp.a);
    const uint64_t col = side_index(key, sp.sh_b, sp.b);
    const double v = raw_value(vals, kind, i);
    if (v == 0.0) continue;
    if (ban_row >= 0 && digit_count(row, sp.a, ban_row) > 1) continue;
    if (ban_col >= 0 && digit_count(col, sp.b, ban_col) > 1) continue;
    if (smem_r) atomicAdd(&sr[row], v);
    else { double* p = slot_insert(rm, row, overflow); if (p) atomicAdd(p, v); }
    if (smem_c) atomicAdd(&sc[col], v);
    else { double* p = slot_insert(cm, col, overflow); if (p) atomicAdd(p, v); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < smem_r; i += blockDim.x)
    if (sr[i] != 0.0) atomicAdd(&rm.sum[i], sr[i]);
  for (int i = threadIdx.x; i < smem_c; i += blockDim.x)
    if (sc[i] != 0.0) atomicAdd(&cm.sum[i], sc[i]);
}

__global__ void table_terms_kernel(const uint64_t* __restrict__ keys, const void* __restrict__ vals, int kind, double divisor,
                                   int64_t num, const __grid_constant__ SplitDev sp, SideMap rm, SideMap cm,
                                   double* __restrict__ partials) {
  __shared__ double sh[kThreads / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num; i += (int64_t)gridDim.x * blockDim.x) {
    const double raw = raw_value(vals, kind, i);
    if (raw == 0.0) continue;
    const uint64_t key = keys[i];
    double c = slot_find(rm, side_index(key, sp.sh_a, sp.a));  // row sum
    double r = slot_find(cm, side_index(key, sp.sh_b, sp.b));  // column sum
    double v = raw;
    if (divisor > 0.0) { c /= divisor; r /= divisor; v /= divisor; }  // counts -> probabilities (fasta.py:66-70)
    acc += v * log(v / (r * c));
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// rows / cols of every pattern plus a flag for the entries the banned-state rule zeroes (constructions.py:94-99)
__global__ void coo_banned_kernel(const uint64_t* __restrict__ keys, int64_t num, const __grid_constant__ SplitDev sp,
                                  int ban_row, int ban_col, int64_t* __restrict__ rows, int64_t* __restrict__ cols,
                                  uint8_t* __restrict__ banned) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num) return;
  const uint64_t key = keys[i];
  const uint64_t row = side_index(key, sp.sh_a, sp.a);
  const uint64_t col = side_index(key, sp.sh_b, sp.b);
  rows[i] = (int64_t)row;
  cols[i] = (int64_t)col;
  banned[i] = (ban_row >= 0 && digit_count(row, sp.a, ban_row) > 1) || (ban_col >= 0 && digit_count(col, sp.b, ban_col) > 1);
}

bool one_side_each(const spb_split* s) {
  uint64_t seen = 0;
  for (int i = 0; i < s->a; ++i) seen |= 1ull << s->idx_a[i];
  for (int i = 0; i < s->b; ++i) seen |= 1ull << s->idx_b[i];
  return s->a + s->b == s->n && seen == (s->n == 64 ? ~0ull : (1ull << s->n) - 1);
}

constexpr int kMaxPartials = 148 * 8;
constexpr int kSmemSide = 2048;  // doubles per side kept in shared memory

inline int reduce_grid(int64_t units) {
  int64_t g = units < 1 ? 1 : units;
  const int64_t cap = (int64_t)sm_count() * 8 < kMaxPartials ? (int64_t)sm_count() * 8 : kMaxPartials;
  return (int)(g < cap ? g : cap);
}

int side_map(const SplitDev& sp, int side_len, double* d_sum, uint64_t* d_hkeys, int64_t cap, SideMap* m, const char* what) {
  m->sum = d_sum;
  m->hkeys = d_hkeys;
  m->mask = 0;
  if (!d_sum) { set_error("%s: NULL sum buffer", what); return SPB_ERR_ARG; }
  if (d_hkeys) {
    if (cap < 2 || (cap & (cap - 1))) { set_error("%s: hash capacity must be a power of two (got %lld)", what, (long long)cap); return SPB_ERR_ARG; }
    m->mask = (uint64_t)cap - 1;
  } else {
    if (side_len > 15) { set_error("%s: a direct-indexed side is limited to 15 taxa; pass a hash table", what); return SPB_ERR_ARG; }
    if (cap < (1ll << (2 * side_len))) { set_error("%s: sum buffer holds %lld entries, 4^%d needed", what, (long long)cap, side_len); return SPB_ERR_ARG; }
  }
  (void)sp;
  return SPB_OK;
}

}  // namespace
}  // namespace spb

using namespace spb;

extern "C" int64_t spb_mi_partials(void) { return kMaxPartials; }

extern "C" int spb_marginals_dense(const double* d_F, int64_t rows, int64_t cols, int64_t ld, double* d_rowsum, double* d_colsum,
                                   void* stream) {
  SPB_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols, "spb_marginals_dense: bad shape %lld x %lld (ld %lld)", (long long)rows,
              (long long)cols, (long long)ld);
  if (rows == 0 || cols == 0) {
    if (rows && d_rowsum) SPB_CUDA(cudaMemsetAsync(d_rowsum, 0, (size_t)rows * sizeof(double), (cudaStream_t)stream));
    if (cols && d_colsum) SPB_CUDA(cudaMemsetAsync(d_colsum, 0, (size_t)cols * sizeof(double), (cudaStream_t)stream));
    return SPB_OK;
  }
  SPB_REQUIRE(d_F && d_rowsum && d_colsum, "spb_marginals_dense: NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  rowsum_kernel<<<(int)(rows < 65535 * 16 ? rows : 65535 * 16), kThreads, 0, st>>>(d_F, rows, cols, ld, d_rowsum);
  SPB_LAUNCH_CHECK();
  colsum_kernel<<<nblk(cols, 32), dim3(32, 8), 0, st>>>(d_F, rows, cols, ld, d_colsum);
  SPB_LAUNCH_CHECK();
  return SPB_OK;
}

extern "C" int spb_outer_f64(const double* d_x, int64_t nx, const double* d_y, int64_t ny, double* d_out, int accumulate,
                             void* stream) {
  SPB_REQUIRE(nx >= 0 && ny >= 0, "spb_outer_f64: negative size");
  if (nx == 0 || ny == 0) return SPB_OK;
  SPB_REQUIRE(d_x && d_y && d_out, "spb_outer_f64: NULL buffer");
  const int64_t total = nx * ny;
  int grid = nblk(total, kThreads);
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  outer_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_x, nx, d_y, ny, d_out, accumulate);
  SPB_LAUNCH_CHECK();
  return SPB_OK;
}

extern "C" int spb_mi_dense(const double* d_F, int64_t rows, int64_t cols, int64_t ld, const double* d_rowsum,
                            const double* d_colsum, double* d_partials, double* d_out, void* stream) {
  SPB_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols, "spb_mi_dense: bad shape");
  SPB_REQUIRE(d_out, "spb_mi_dense: NULL output");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0 || cols == 0) {
    SPB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
    return SPB_OK;
  }
  SPB_REQUIRE(d_F && d_rowsum && d_colsum && d_partials, "spb_mi_dense: NULL buffer");
  const int grid = reduce_grid(rows);
  dense_terms_kernel<<<grid, kThreads, 0, st>>>(d_F, rows, cols, ld, d_rowsum, d_colsum, d_partials);
  SPB_LAUNCH_CHECK();
  sum_partials_kernel<<<1, kThreads, 0, st>>>(d_partials, grid, d_out);
  SPB_LAUNCH_CHECK();
  return SPB_OK;
}

extern "C" int spb_flatten_coo_banned(const uint64_t* d_keys, int64_t num, const spb_split* split, int ban_row, int ban_col,
                                      int64_t* d_rows, int64_t* d_cols, uint8_t* d_banned, void* stream) {
  SplitDev sp;
  int rc = make_split_dev(split, &sp);
  if (rc) return rc;
  SPB_REQUIRE(sp.a <= 31 && sp.b <= 31, "spb_flatten_coo_banned: sides are limited to 31 taxa (int64 indices)");
  SPB_REQUIRE(ban_row >= -1 && ban_row <= 3 && ban_col >= -1 && ban_col <= 3, "spb_flatten_coo_banned: banned state must be -1 or 0..3");
  if (num <= 0) return SPB_OK;
  SPB_REQUIRE(d_keys && d_rows && d_cols && d_banned, "spb_flatten_coo_banned: NULL buffer");
  coo_banned_kernel<<<nblk(num, kThreads), kThreads, 0, (cudaStream_t)stream>>>(d_keys, num, sp, ban_row, ban_col, d_rows, d_cols, d_banned);
  SPB_LAUNCH_CHECK();
  return SPB_OK;
}

extern "C" int spb_table_marginals(const uint64_t* d_keys, const void* d_vals, int val_kind, int64_t num, const spb_split* split,
                                   int ban_row, int ban_col, double* d_rowsum, uint64_t* d_rkeys, int64_t rcap,
                                   double* d_colsum, uint64_t* d_ckeys, int64_t ccap, uint32_t* d_overflow, void* stream) {
  SplitDev sp;
  int rc = make_split_dev(split, &sp);
  if (rc) return rc;
  SPB_REQUIRE(sp.a <= 31 && sp.b <= 31, "spb_table_marginals: sides are limited to 31 taxa");
  SPB_REQUIRE(ban_row >= -1 && ban_row <= 3 && ban_col >= -1 && ban_col <= 3, "spb_table_marginals: banned state must be -1 or 0..3");
  SPB_REQUIRE(val_kind == SPB_VAL_U32 || val_kind == SPB_VAL_F64, "spb_table_marginals: bad value kind");
  if (!one_side_each(split)) {
    set_error("spb_table_marginals: the split must place every taxon on exactly one side (materialise the flattening otherwise)");
    return SPB_ERR_UNSUPPORTED;
  }
  SideMap rm, cm;
  if ((rc = side_map(sp, sp.a, d_rowsum, d_rkeys, rcap, &rm, "spb_table_marginals(rows)"))) return rc;
  if ((rc = side_map(sp, sp.b, d_colsum, d_ckeys, ccap, &cm, "spb_table_marginals(cols)"))) return rc;
  if (num <= 0) return SPB_OK;
  SPB_REQUIRE(d_keys && d_vals && d_overflow, "spb_table_marginals: NULL buffer");
  const int smem_r = (!d_rkeys && (1ll << (2 * sp.a)) <= kSmemSide) ? (int)(1ll << (2 * sp.a)) : 0;
  const int smem_c = (!d_ckeys && (1ll << (2 * sp.b)) <= kSmemSide) ? (int)(1ll << (2 * sp.b)) : 0;
  int grid = nblk(num, kThreads);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  table_marginals_kernel<<<grid, kThreads, (size_t)(smem_r + smem_c) * sizeof(double), (cudaStream_t)stream>>>(
      d_keys, d_vals, val_kind, num, sp, ban_row, ban_col, rm, cm, smem_r, smem_c, d_overflow);
  SPB_LAUNCH_CHECK();
  return SPB_OK;
}

extern "C" int spb_mi_table(const uint64_t* d_keys, const void* d_vals, int val_kind, double divisor, int64_t num,
                            const spb_split* split, const double* d_rowsum, const uint64_t* d_rkeys, int64_t rcap,
                            const double* d_colsum, const uint64_t* d_ckeys, int64_t ccap, double* d_partials, double* d_out,
                            void* stream) {
  SplitDev sp;
  int rc = make_split_dev(split, &sp);
  if (rc) return rc;
  SPB_REQUIRE(sp.a <= 31 && sp.b <= 31, "spb_mi_table: sides are limited to 31 taxa");
  SPB_REQUIRE(val_kind == SPB_VAL_U32 || val_kind == SPB_VAL_F64, "spb_mi_table: bad value kind");
  SPB_REQUIRE(d_out, "spb_mi_table: NULL output");
  if (!one_side_each(split)) {
    set_error("spb_mi_table: the split must place every taxon on exactly one side (materialise the flattening otherwise)");
    return SPB_ERR_UNSUPPORTED;
  }
  SideMap rm, cm;
  if ((rc = side_map(sp, sp.a, const_cast<double*>(d_rowsum), const_cast<uint64_t*>(d_rkeys), rcap, &rm, "spb_mi_table(rows)"))) return rc;
  if ((rc = side_map(sp, sp.b, const_cast<double*>(d_colsum), const_cast<uint64_t*>(d_ckeys), ccap, &cm, "spb_mi_table(cols)"))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (num <= 0) {
    SPB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
    return SPB_OK;
  }
  SPB_REQUIRE(d_keys && d_vals && d_partials, "spb_mi_table: NULL buffer");
  const int grid = reduce_grid(nblk(num, kThreads));
  table_terms_kernel<<<grid, kThreads, 0, st>>>(d_keys, d_vals, val_kind, divisor, num, sp, rm, cm, d_partials);
  SPB_LAUNCH_CHECK();
  sum_partials_kernel<<<1, kThreads, 0, st>>>(d_partials, grid, d_out);
  SPB_LAUNCH_CHECK();
  return SPB_OK;
}
