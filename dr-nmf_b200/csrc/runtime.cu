// Host-side runtime bits of libdrnmf.so: error strings, launch counter, TMA tensor-map construction.
#include "internal.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cudaTypedefs.h>

namespace drnmf {

static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int stream_wait_geq(cudaStream_t st, const unsigned int* addr, unsigned int value) {
  static PFN_cuStreamWaitValue32_v11070 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuStreamWaitValue32_v11070>(p);
  }
  if (!fn) return 1;
  if (!addr) return 0;                                   // availability probe
  return fn(st, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS ? 0 : 1;
}

int make_tmap_2d(CUtensorMap* out, const float* base, uint64_t cols, uint64_t rows, uint64_t row_stride_elems,
                 uint32_t box_cols, uint32_t box_rows) {
  auto enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return DRNMF_ERR_CUDA; }
  DRNMF_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  DRNMF_CHECK((row_stride_elems * sizeof(float)) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  DRNMF_CHECK(box_cols * sizeof(float) == 128 && box_rows <= 256, "TMA box must be 128 bytes wide, <= 256 rows");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {row_stride_elems * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu stride=%llu box=%ux%u)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_elems, box_cols,
              box_rows);
    return DRNMF_ERR_CUDA;
  }
  return DRNMF_OK;
}

// fp32 row-major matrix (rows x cols, cols a multiple of 32) seen as [slab = cols/32][row][32 columns]: one TMA box of
// (32 columns x box_rows x box_slabs) lands in shared memory as box_slabs consecutive K-major tiles of box_rows x 128
// bytes, each in the 128B-swizzle layout the tcgen05 descriptors expect (a 2-D map moves one such tile per instruction).
int make_tmap_slabs(CUtensorMap* out, const float* base, uint64_t cols, uint64_t rows, uint64_t row_stride_elems,
                    uint32_t box_rows, uint32_t box_slabs) {
  auto enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return DRNMF_ERR_CUDA; }
  DRNMF_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  DRNMF_CHECK((row_stride_elems * sizeof(float)) % 16 == 0 && cols % 32 == 0, "TMA slab map needs cols %% 32 == 0");
  DRNMF_CHECK(box_rows <= 256 && box_slabs >= 1 && box_slabs <= 256, "TMA box out of range");
  cuuint64_t gdim[3] = {32, rows, cols / 32};
  cuuint64_t gstr[2] = {row_stride_elems * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, box_rows, box_slabs};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (slabs) failed with CUresult %d (cols=%llu rows=%llu stride=%llu box=32x%ux%u)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_elems, box_rows, box_slabs);
    return DRNMF_ERR_CUDA;
  }
  return DRNMF_OK;
}

}  // namespace drnmf
