// Host-side helpers shared by the tcgen05 translation units (included inside namespace d2r::<anon>): tensor-map
// encoding through the driver entry point, programmatic-dependent-launch wrapper, device index.
#pragma once

constexpr int kTcThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &f, 12000, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(f);
    tried = true;
  }
  return fn;
}

// 4-D map {inner (contiguous), rows, batch_inner, batch_outer}; bf16; 128-byte swizzle; OOB -> 0
inline int encode_map(CUtensorMap* tm, const void* base, int es, long long inner, long long rows, long long bi, long long bo,
               long long ld, long long si, long long so, int box_inner, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(D2R_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const long long row_bytes = ld * es;
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)bi, (cuuint64_t)bo};
  cuuint64_t strides[3] = {(cuuint64_t)row_bytes, (cuuint64_t)(bi > 1 ? si * es : row_bytes),
                           (cuuint64_t)(bo > 1 ? so * es : row_bytes)};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(D2R_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): dims=[%lld,%lld,%lld,%lld] ld=%lld si=%lld so=%lld", (int)r,
                     inner, rows, bi, bo, ld, si, so);
  return D2R_OK;
}

inline int encode_operand(CUtensorMap* tm, const void* base, long long inner, long long rows, long long bi, long long bo,
                   long long ld, long long si, long long so, int box_rows) {
  return encode_map(tm, base, 2, inner, rows, bi, bo, ld, si, so, 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may be scheduled while its
// predecessor on the stream is still draining; it orders itself with griddepcontrol.wait after its prologue.
// D2R_PDL=0 in the environment launches normally.
template <typename Kern, typename... Args>
cudaError_t launch_pdl(Kern kern, dim3 grid, int smem_bytes, cudaStream_t stream, const Args&... args) {
  static int use_pdl = -1;
  if (use_pdl < 0) {
    const char* e = getenv("D2R_PDL");
    use_pdl = (e && atoi(e) == 0) ? 0 : 1;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

