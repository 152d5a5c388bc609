// store.cu -- the HBM-resident corpus store (K0 ingest, tombstones) and library-wide plumbing.
//
// Replaces the Chroma collection the reference creates at rag_engine/storage/vector_store.py:44-52
// and writes through rag_engine/storage/vector_store.py:68-82 (collection.add) / :102-105 (delete).
//
// HBM layout (row-major, one allocation per array, sized for `capacity_rows` up front so device
// pointers and the TMA descriptor never move):
//   f32   [cap, D]  raw rows as appended                       (CMW_STORE_F32;  exact rescoring, K1)
//   bf16  [cap, D]  rows divided by their fp64 norm, RN to bf16 (CMW_STORE_BF16; K1 bf16, K2)
//   inv_norm/norm/live f32[cap]   per-row multipliers, NaN once tombstoned (a NaN score fails every
//                                 `score >= thr` test, so dead rows are never admitted)
//   norm64 f64[cap], kb_gid i32[cap]
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace cmw {

static thread_local std::string t_last_error;
std::atomic<long long> g_kernel_launches{0};
std::atomic<int> g_pdl{1};  // programmatic dependent launch of the small-batch kernel chain (option "pdl")
thread_local int t_pdl_search = 0;  // set per search by run_filter_half: this search's kernels are launched that way
Options g_opt;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
}

// ---------------------------------------------------------------------------------------------
// K0: ingest.  One warp per row: fp64 norm (fixed summation order), raw fp32 copy, normalised bf16.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ingest_kernel(const float* __restrict__ src, const int32_t* __restrict__ gid_src,
              const int64_t* __restrict__ src_index, int64_t n, int dim, int64_t row0, float* __restrict__ f32, __nv_bfloat16* __restrict__ bf16, int half_tiles,
              float* __restrict__ inv_norm, float* __restrict__ norm, float* __restrict__ live,
              double* __restrict__ norm64, int32_t* __restrict__ kb_gid,
              uint32_t* __restrict__ maxnorm_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim >> 2;
    for (int64_t r = warp; r < n; r += nwarps) {
        // (device-to-device re-ingest gathers its source rows through an index; plain appends read row r)
        const int64_t sr = src_index != nullptr ? src_index[r] : r;
        const float4* in = reinterpret_cast<const float4*>(src + sr * dim);
        double acc = 0.0, t32 = 0.0;  // t32: |c - c with its low 13 mantissa bits cleared|^2 (what a tf32 MMA drops)
        for (int c = lane; c < nvec; c += 32) {
            float4 v = __ldg(in + c);
            acc += (double)v.x * (double)v.x;
            acc += (double)v.y * (double)v.y;
            acc += (double)v.z * (double)v.z;
            acc += (double)v.w * (double)v.w;
            const double dx = (double)v.x - (double)__uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            const double dy = (double)v.y - (double)__uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            const double dz = (double)v.z - (double)__uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            const double dw = (double)v.w - (double)__uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            t32 += dx * dx + dy * dy + dz * dz + dw * dw;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            t32 += __shfl_xor_sync(0xffffffffu, t32, o);
        }
        const double nrm = sqrt(acc);
        const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
        const int64_t dst = row0 + r;
        if (f32 != nullptr) {
            float4* out = reinterpret_cast<float4*>(f32 + dst * dim);
            for (int c = lane; c < nvec; c += 32) out[c] = __ldg(in + c);
        }
        double res2 = 0.0;  // |c/|c| - its bf16 tile|^2 from the values actually written (rigorous certificate)
        if (bf16 != nullptr) {
            uint2* out = reinterpret_cast<uint2*>(bf16 + dst * dim);
            for (int c = lane; c < nvec; c += 32) {
                float4 v = __ldg(in + c);
                const double ex = (double)v.x * inv, ey = (double)v.y * inv, ez = (double)v.z * inv,
                             ew = (double)v.w * inv;
                float sx, sy, sz, sw;
                const uint32_t tx = to_tile16((float)ex, half_tiles, sx), ty = to_tile16((float)ey, half_tiles, sy);
                const uint32_t tz = to_tile16((float)ez, half_tiles, sz), tw = to_tile16((float)ew, half_tiles, sw);
                out[c] = make_uint2(tx | (ty << 16), tz | (tw << 16));
                const double dx = ex - (double)sx, dy = ey - (double)sy, dz = ez - (double)sz, dw = ew - (double)sw;
                res2 += dx * dx + dy * dy + dz * dz + dw * dw;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) res2 += __shfl_xor_sync(0xffffffffu, res2, o);
        }
        // 4-norm of the normalised row (certificate bound of the bf16 filter, Options::bf16_eps)
        double s4 = 0.0;
        for (int c = lane; c < nvec; c += 32) {
            float4 v = __ldg(in + c);
            const double x = (double)v.x * inv, y = (double)v.y * inv, z = (double)v.z * inv, w = (double)v.w * inv;
            s4 += x * x * x * x + y * y * y * y + z * z * z * z + w * w * w * w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s4 += __shfl_xor_sync(0xffffffffu, s4, o);
        if (lane == 0) {
            atomicMax(maxnorm_bits + 1, __float_as_uint((float)sqrt(sqrt(s4))) + 1u);
            // rounded up, and a hair more for the fp64 arithmetic above (positive floats order like their bits)
            atomicMax(maxnorm_bits + 2, __float_as_uint(__double2float_ru(sqrt(res2) * (1.0 + 1e-9) + 1e-12)));
            // truncation bounds every round-to-nearest variant element by element, so this holds whichever way
            // the tensor core narrows fp32 to tf32
            atomicMax(maxnorm_bits + 3, __float_as_uint(__double2float_ru(sqrt(t32) * inv * (1.0 + 1e-9) + 1e-12)));
            inv_norm[dst] = (float)inv;
            norm[dst] = (float)nrm;
            live[dst] = 1.0f;
            norm64[dst] = nrm;
            kb_gid[dst] = gid_src != nullptr ? gid_src[sr] : -1;
            atomicMax(maxnorm_bits, __float_as_uint((float)nrm) + 1u);  // +1 ulp: upper bound
        }
    }
}

__global__ void tombstone_kernel(const int64_t* __restrict__ rows, int64_t n, int64_t nrows,
                                 float* inv_norm, float* norm, float* live,
                                 unsigned long long* newly_dead, uint32_t* dead_blk) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t r = rows[i];
    if (r < 0 || r >= nrows) return;
    const float qnan = __int_as_float(0x7fc00000);
    // atomic exchange on `live` so that duplicates in `rows` are counted once
    float old = __uint_as_float(atomicExch(reinterpret_cast<unsigned int*>(live + r), 0x7fc00000u));
    if (old == old) {
        inv_norm[r] = qnan;
        norm[r] = qnan;
        atomicAdd(newly_dead, 1ull);
        atomicAdd(dead_blk + (r >> 8), 1u);
    }
}

static int ensure_stream(Store* s) {
    if (s->stream == nullptr) CMW_CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    return 0;
}

int ensure_pinned(Store* s, size_t bytes) {
    if (s->pinned_bytes >= bytes) return 0;
    if (s->pinned) cudaFreeHost(s->pinned);
    s->pinned = nullptr;
    s->pinned_bytes = 0;
    CMW_CUDA_OK(cudaMallocHost(&s->pinned, bytes));
    s->pinned_bytes = bytes;
    return 0;
}

int ensure_dev_io(Store* s, size_t bytes) {
    if (s->dev_io_bytes >= bytes) return 0;
    if (s->dev_io) cudaFree(s->dev_io);
    s->dev_io = nullptr;
    s->dev_io_bytes = 0;
    CMW_CUDA_OK(cudaMalloc(&s->dev_io, bytes));
    s->dev_io_bytes = bytes;
    return 0;
}

int ensure_ws(Store* s, size_t bytes) {
    if (s->ws_bytes >= bytes) return 0;
    if (s->ws) cudaFree(s->ws);
    s->ws = nullptr;
    s->ws_bytes = 0;
    CMW_CUDA_OK(cudaMalloc(&s->ws, bytes));
    s->ws_bytes = bytes;
    return 0;
}

int get_stream(Store* s, cudaStream_t* out) {
    int rc = ensure_stream(s);
    if (rc) return rc;
    *out = s->stream;
    return 0;
}

int encode_bf16_tmap(Store* s);  // gemm.cu
int encode_f32_tmap(Store* s);   // gemm.cu

// ---------------------------------------------------------------------------------------------
// host staging for the *_host entry points
// ---------------------------------------------------------------------------------------------
int stage_h2d(int device, uint8_t* pin, uint8_t* dev, const uint8_t* src, size_t bytes,
                     cudaStream_t stream) {
    const size_t piece = 1u << 20;
    const size_t npieces = (bytes + piece - 1) / piece;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nthreads = hw >= 8 ? 4 : (hw >= 4 ? 2 : 1);
    if (npieces < 4) nthreads = 1;
    std::atomic<int> err{0};
    auto work = [&](size_t t) {
        if (t != 0 && cudaSetDevice(device) != cudaSuccess) err.store(1);
        for (size_t i = t; i < npieces; i += nthreads) {
            const size_t off = i * piece;
            const size_t n = bytes - off < piece ? bytes - off : piece;
            memcpy(pin + off, src + off, n);
            if (cudaMemcpyAsync(dev + off, pin + off, n, cudaMemcpyHostToDevice, stream) != cudaSuccess)
                err.store(1);
        }
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t t = 1; t < nthreads; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    if (err.load()) {
        set_error("host staging copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -2;
    }
    return 0;
}


}  // namespace cmw

using namespace cmw;

extern "C" {

const char* cmw_last_error(void) { return t_last_error.c_str(); }
int cmw_abi_version(void) { return CMW_ABI_VERSION; }
int64_t cmw_kernel_launches(void) { return (int64_t)g_kernel_launches.load(); }

int cmw_set_option(const char* name, double value) {
    if (!name) return -1;
    if (!strcmp(name, "bf16_eps")) g_opt.bf16_eps = value;
    else if (!strcmp(name, "f32_eps")) g_opt.f32_eps = value;
    else if (!strcmp(name, "bf16_sigmas")) g_opt.bf16_sigmas = value;
    else if (!strcmp(name, "kprime")) g_opt.kprime = value;
    else if (!strcmp(name, "scan_max_batch")) g_opt.scan_max_batch = value;
    else if (!strcmp(name, "gemm_enabled")) g_opt.gemm_enabled = value;
    else if (!strcmp(name, "slab_growth")) g_opt.slab_growth = value;
    else if (!strcmp(name, "strict_certificate")) g_opt.strict_certificate = value;
    else if (!strcmp(name, "repair")) g_opt.repair = value;
    else if (!strcmp(name, "host_overlap")) g_opt.host_overlap = value;
    else if (!strcmp(name, "wide_dense")) g_opt.wide_dense = value;
    else if (!strcmp(name, "pdl")) g_pdl.store(value != 0 ? 1 : 0);
    else if (!strcmp(name, "scan_permute")) g_opt.scan_permute = value;
    else if (!strcmp(name, "gemm_2cta")) g_opt.gemm_2cta = value;
    else if (!strcmp(name, "gemm_clc")) g_opt.gemm_clc = value;
    else if (!strcmp(name, "gemm_2cta_min_batch")) g_opt.gemm_2cta_min_batch = value;
    else if (!strcmp(name, "f16_bits")) g_opt.f16_bits = value;
    else {
        set_error("unknown option '%s'", name);
        return -1;
    }
    return 0;
}

double cmw_get_option(const char* name) {
    if (!name) return 0.0;
    if (!strcmp(name, "bf16_eps")) return g_opt.bf16_eps;
    if (!strcmp(name, "f32_eps")) return g_opt.f32_eps;
    if (!strcmp(name, "bf16_sigmas")) return g_opt.bf16_sigmas;
    if (!strcmp(name, "kprime")) return g_opt.kprime;
    if (!strcmp(name, "scan_max_batch")) return g_opt.scan_max_batch;
    if (!strcmp(name, "gemm_enabled")) return g_opt.gemm_enabled;
    if (!strcmp(name, "slab_growth")) return g_opt.slab_growth;
    if (!strcmp(name, "strict_certificate")) return g_opt.strict_certificate;
    if (!strcmp(name, "repair")) return g_opt.repair;
    if (!strcmp(name, "host_overlap")) return g_opt.host_overlap;
    if (!strcmp(name, "wide_dense")) return g_opt.wide_dense;
    if (!strcmp(name, "pdl")) return (double)g_pdl.load();
    if (!strcmp(name, "scan_permute")) return g_opt.scan_permute;
    if (!strcmp(name, "gemm_2cta")) return g_opt.gemm_2cta;
    if (!strcmp(name, "gemm_clc")) return g_opt.gemm_clc;
    if (!strcmp(name, "gemm_2cta_min_batch")) return g_opt.gemm_2cta_min_batch;
    if (!strcmp(name, "f16_bits")) return g_opt.f16_bits;
    if (!strcmp(name, "pool_cap")) return (double)kPoolCap;
    return 0.0;
}

int cmw_store_create(int device, int dim, int64_t capacity_rows, uint32_t flags, int64_t id_offset,
                     cmw_store** out) {
    CMW_REQUIRE(out != nullptr, "cmw_store_create: out is NULL");
    *out = nullptr;
    CMW_REQUIRE(dim >= 8 && dim % 8 == 0 && dim <= 8192,
                "cmw_store_create: dim must be a multiple of 8 in [8, 8192], got %d", dim);
    CMW_REQUIRE(capacity_rows > 0 && capacity_rows < (1ll << 31),
                "cmw_store_create: capacity_rows must be in (0, 2^31), got %lld",
                (long long)capacity_rows);
    CMW_REQUIRE((flags & (CMW_STORE_F32 | CMW_STORE_BF16 | CMW_STORE_F16)) != 0,
                "cmw_store_create: flags must include CMW_STORE_F32 and/or one of CMW_STORE_BF16 / CMW_STORE_F16");
    CMW_REQUIRE((flags & (CMW_STORE_BF16 | CMW_STORE_F16)) != (CMW_STORE_BF16 | CMW_STORE_F16),
                "cmw_store_create: CMW_STORE_BF16 and CMW_STORE_F16 are alternatives (one set of 16-bit tiles)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        set_error("cmw_store_create: no CUDA device (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return -3;
    }
    CMW_REQUIRE(device >= 0 && device < ndev, "cmw_store_create: device %d out of range (0..%d)",
                device, ndev - 1);
    CMW_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CMW_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    CMW_REQUIRE(prop.major == 10, "cmw_store_create: device %d is sm_%d%d; this library is sm_100a only",
                device, prop.major, prop.minor);
    Store* s = new Store();
    s->device = device;
    s->dim = dim;
    s->flags = flags;
    s->sm_count = prop.multiProcessorCount;
    s->capacity = capacity_rows;
    s->id_offset = id_offset;
    const int64_t cap4 = ((capacity_rows + 3) / 4) * 4 + 4;
    auto alloc = [&](void** p, size_t bytes) -> int {
        CMW_CUDA_OK(cudaMalloc(p, bytes));
        s->hbm_bytes += bytes;
        return 0;
    };
    int rc = 0;
    const size_t elems = (size_t)capacity_rows * (size_t)dim;
    if (!rc && (flags & CMW_STORE_F32)) rc = alloc((void**)&s->f32, elems * sizeof(float));
    s->half_tiles = (flags & CMW_STORE_F16) != 0;
    s->half_bits = (int)g_opt.f16_bits < 8 ? 8 : ((int)g_opt.f16_bits > 11 ? 11 : (int)g_opt.f16_bits);
    if (!rc && (flags & (CMW_STORE_BF16 | CMW_STORE_F16))) rc = alloc((void**)&s->bf16, elems * sizeof(__nv_bfloat16));
    if (!rc) rc = alloc((void**)&s->inv_norm, cap4 * sizeof(float));
    if (!rc) rc = alloc((void**)&s->norm, cap4 * sizeof(float));
    if (!rc) rc = alloc((void**)&s->live, cap4 * sizeof(float));
    if (!rc) rc = alloc((void**)&s->norm64, capacity_rows * sizeof(double));
    if (!rc) rc = alloc((void**)&s->kb_gid, capacity_rows * sizeof(int32_t));
    if (!rc) rc = alloc((void**)&s->maxnorm_bits, 256);
    const size_t dead_blk_bytes = (size_t)(capacity_rows / 256 + 1) * sizeof(uint32_t);
    if (!rc) rc = alloc((void**)&s->dead_blk, dead_blk_bytes);
    if (!rc) rc = ensure_stream(s);
    if (!rc) {
        cudaError_t e2 = cudaMemsetAsync(s->maxnorm_bits, 0, 256, s->stream);
        if (e2 == cudaSuccess) e2 = cudaMemsetAsync(s->dead_blk, 0, dead_blk_bytes, s->stream);
        if (e2 == cudaSuccess) e2 = cudaMemsetAsync(s->inv_norm, 0xff, cap4 * sizeof(float), s->stream);
        if (e2 == cudaSuccess) e2 = cudaMemsetAsync(s->norm, 0xff, cap4 * sizeof(float), s->stream);
        if (e2 == cudaSuccess) e2 = cudaMemsetAsync(s->live, 0xff, cap4 * sizeof(float), s->stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(s->stream);
        if (e2 != cudaSuccess) {
            set_error("cmw_store_create: init failed: %s", cudaGetErrorString(e2));
            rc = -2;
        }
    }
    if (!rc && s->bf16 != nullptr) {
        // a failed encode only disables K2 (the store stays usable through K1); reported by info
        s->tmap_ok = (encode_bf16_tmap(s) == 0);
    }
    if (!rc && s->f32 != nullptr) s->tmap_f32_ok = (encode_f32_tmap(s) == 0);
    if (rc) {
        cmw_store_destroy(reinterpret_cast<cmw_store*>(s));
        return rc;
    }
    *out = reinterpret_cast<cmw_store*>(s);
    return 0;
}

int cmw_store_destroy(cmw_store* h) {
    if (!h) return 0;
    Store* s = reinterpret_cast<Store*>(h);
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_in) cudaStreamSynchronize(s->copy_in);
    if (s->tail) cudaStreamSynchronize(s->tail);
    if (s->copy_out) cudaStreamSynchronize(s->copy_out);
    for (HostSlot& sl : s->slots) {
        cudaFree(sl.dev_io);
        cudaFree(sl.ws);
        if (sl.pinned) cudaFreeHost(sl.pinned);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
        if (sl.ev_compute) cudaEventDestroy(sl.ev_compute);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (s->copy_in) cudaStreamDestroy(s->copy_in);
    if (s->copy_out) cudaStreamDestroy(s->copy_out);
    if (s->tail) cudaStreamDestroy(s->tail);
    cudaFree(s->f32);
    cudaFree(s->bf16);
    cudaFree(s->inv_norm);
    cudaFree(s->norm);
    cudaFree(s->live);
    cudaFree(s->norm64);
    cudaFree(s->kb_gid);
    cudaFree(s->maxnorm_bits);
    cudaFree(s->dead_blk);
    cudaFree(s->dev_io);
    cudaFree(s->ws);
    if (s->pinned) cudaFreeHost(s->pinned);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

int cmw_store_get_info(const cmw_store* h, cmw_store_info* out) {
    CMW_REQUIRE(h && out, "cmw_store_get_info: NULL argument");
    const Store* s = reinterpret_cast<const Store*>(h);
    out->device = s->device;
    out->dim = s->dim;
    out->flags = s->flags | ((s->tmap_ok || s->tmap_f32_ok) ? 0x100u : 0u) | (s->tmap_f32_ok ? 0x200u : 0u);
    out->sm_count = s->sm_count;
    out->capacity_rows = s->capacity;
    out->rows = s->rows;
    out->live_rows = s->rows - s->dead;
    out->id_offset = s->id_offset;
    out->hbm_bytes = (int64_t)s->hbm_bytes;
    return 0;
}

int cmw_store_append_f32(cmw_store* h, const float* rows_dev, const int32_t* kb_gid_dev, int64_t n,
                         void* stream) {
    CMW_REQUIRE(h != nullptr, "cmw_store_append_f32: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    if (n == 0) return 0;
    CMW_REQUIRE(n > 0 && rows_dev != nullptr, "cmw_store_append_f32: bad arguments");
    CMW_REQUIRE(s->rows + n <= s->capacity,
                "cmw_store_append_f32: %lld rows + %lld exceed the capacity %lld", (long long)s->rows,
                (long long)n, (long long)s->capacity);
    CMW_REQUIRE((reinterpret_cast<uintptr_t>(rows_dev) & 15) == 0,
                "cmw_store_append_f32: rows_dev must be 16-byte aligned");
    CMW_CUDA_OK(cudaSetDevice(s->device));
    const int warps_per_block = 8;
    int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
    const int64_t max_blocks = (int64_t)s->sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    ingest_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, (cudaStream_t)stream>>>(
        rows_dev, kb_gid_dev, nullptr, n, s->dim, s->rows, s->f32, s->bf16, s->half_tiles ? s->half_bits : 0, s->inv_norm, s->norm, s->live,
        s->norm64, s->kb_gid, s->maxnorm_bits);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    s->rows += n;
    return 0;
}

int cmw_store_copy_rows(cmw_store* dst_h, const cmw_store* src_h, const int64_t* src_rows_dev, int64_t src_row0,
                        int64_t n, void* stream) {
    CMW_REQUIRE(dst_h != nullptr && src_h != nullptr, "cmw_store_copy_rows: store is NULL");
    Store* d = reinterpret_cast<Store*>(dst_h);
    const Store* s = reinterpret_cast<const Store*>(src_h);
    if (n == 0) return 0;
    CMW_REQUIRE(n > 0, "cmw_store_copy_rows: bad row count");
    CMW_REQUIRE(d != s, "cmw_store_copy_rows: source and destination are the same store");
    CMW_REQUIRE(s->f32 != nullptr, "cmw_store_copy_rows: the source store keeps no fp32 tiles to re-ingest from");
    CMW_REQUIRE(s->device == d->device && s->dim == d->dim,
                "cmw_store_copy_rows: stores must live on the same device and have the same dim");
    CMW_REQUIRE(d->rows + n <= d->capacity, "cmw_store_copy_rows: %lld rows + %lld exceed the capacity %lld",
                (long long)d->rows, (long long)n, (long long)d->capacity);
    if (src_rows_dev == nullptr)
        CMW_REQUIRE(src_row0 >= 0 && src_row0 + n <= s->rows, "cmw_store_copy_rows: rows [%lld, %lld) out of range",
                    (long long)src_row0, (long long)(src_row0 + n));
    CMW_CUDA_OK(cudaSetDevice(d->device));
    const int warps_per_block = 8;
    int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
    const int64_t max_blocks = (int64_t)d->sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    const float* base = src_rows_dev ? s->f32 : s->f32 + (size_t)src_row0 * s->dim;
    const int32_t* gid = src_rows_dev ? s->kb_gid : s->kb_gid + src_row0;
    ingest_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, (cudaStream_t)stream>>>(
        base, gid, src_rows_dev, n, d->dim, d->rows, d->f32, d->bf16, d->half_tiles ? d->half_bits : 0, d->inv_norm, d->norm,
        d->live, d->norm64, d->kb_gid, d->maxnorm_bits);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    d->rows += n;
    return 0;
}

int cmw_store_append_host_f32(cmw_store* h, const float* rows_host, const int32_t* kb_gid_host,
                              int64_t n) {
    CMW_REQUIRE(h != nullptr, "cmw_store_append_host_f32: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    if (n == 0) return 0;
    CMW_REQUIRE(n > 0 && rows_host != nullptr, "cmw_store_append_host_f32: bad arguments");
    CMW_REQUIRE(s->rows + n <= s->capacity,
                "cmw_store_append_host_f32: %lld rows + %lld exceed the capacity %lld",
                (long long)s->rows, (long long)n, (long long)s->capacity);
    std::lock_guard<std::mutex> host_lock(s->host_mu);
    CMW_CUDA_OK(cudaSetDevice(s->device));
    int rc = ensure_stream(s);
    if (rc) return rc;
    const size_t row_bytes = (size_t)s->dim * sizeof(float);
    int64_t chunk = (int64_t)((64u << 20) / row_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    const size_t gid_off = ((size_t)chunk * row_bytes + 255) & ~(size_t)255;
    const size_t need = gid_off + (size_t)chunk * sizeof(int32_t);
    if ((rc = ensure_pinned(s, need))) return rc;
    if ((rc = ensure_dev_io(s, need))) return rc;
    // page-locked source rows are DMA-copied directly; pageable ones go through the pinned block in 1 MB
    // pieces staged by a few host threads, each piece's H2D queued as soon as it is staged
    cudaPointerAttributes attr;
    bool src_pinned = cudaPointerGetAttributes(&attr, rows_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    for (int64_t lo = 0; lo < n; lo += chunk) {
        const int64_t m = (n - lo < chunk) ? (n - lo) : chunk;
        const float* src = rows_host + lo * s->dim;
        if (src_pinned) {
            CMW_CUDA_OK(cudaMemcpyAsync(s->dev_io, src, (size_t)m * row_bytes, cudaMemcpyHostToDevice, s->stream));
        } else if (stage_h2d(s->device, reinterpret_cast<uint8_t*>(s->pinned), reinterpret_cast<uint8_t*>(s->dev_io),
                             reinterpret_cast<const uint8_t*>(src), (size_t)m * row_bytes, s->stream)) {
            return -2;
        }
        if (kb_gid_host) memcpy((char*)s->pinned + gid_off, kb_gid_host + lo, (size_t)m * sizeof(int32_t));
        if (kb_gid_host)
            CMW_CUDA_OK(cudaMemcpyAsync((char*)s->dev_io + gid_off, (char*)s->pinned + gid_off,
                                        (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
        rc = cmw_store_append_f32(h, (const float*)s->dev_io,
                                  kb_gid_host ? (const int32_t*)((char*)s->dev_io + gid_off) : nullptr,
                                  m, s->stream);
        if (rc) return rc;
        CMW_CUDA_OK(cudaStreamSynchronize(s->stream));
    }
    return 0;
}

int cmw_store_tombstone(cmw_store* h, const int64_t* rows_dev, int64_t n, void* stream) {
    CMW_REQUIRE(h != nullptr, "cmw_store_tombstone: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    if (n == 0) return 0;
    CMW_REQUIRE(n > 0 && rows_dev != nullptr, "cmw_store_tombstone: bad arguments");
    CMW_CUDA_OK(cudaSetDevice(s->device));
    // the live-row count is host state: count newly dead rows through a device counter
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(s->maxnorm_bits + 16);
    cudaStream_t st = (cudaStream_t)stream;
    CMW_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    tombstone_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows_dev, n, s->rows, s->inv_norm,
                                                                 s->norm, s->live, counter, s->dead_blk);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    unsigned long long host_count = 0;
    CMW_CUDA_OK(cudaMemcpyAsync(&host_count, counter, sizeof(host_count), cudaMemcpyDeviceToHost, st));
    // where the dead rows are: per-block counts -> host prefix sums for the slab schedule (api.cu)
    const size_t nblk = (size_t)(s->rows / 256 + 1);
    std::vector<uint32_t> blk(nblk);
    CMW_CUDA_OK(cudaMemcpyAsync(blk.data(), s->dead_blk, nblk * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CMW_CUDA_OK(cudaStreamSynchronize(st));
    s->dead += (int64_t)host_count;
    if (s->dead > 0) {
        s->dead_prefix.assign(nblk + 1, 0);
        for (size_t i = 0; i < nblk; ++i) s->dead_prefix[i + 1] = s->dead_prefix[i] + (int64_t)blk[i];
    }
    return 0;
}

int cmw_store_tombstone_host(cmw_store* h, const int64_t* rows_host, int64_t n) {
    CMW_REQUIRE(h != nullptr, "cmw_store_tombstone_host: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    if (n == 0) return 0;
    CMW_REQUIRE(n > 0 && rows_host != nullptr, "cmw_store_tombstone_host: bad arguments");
    std::lock_guard<std::mutex> host_lock(s->host_mu);
    CMW_CUDA_OK(cudaSetDevice(s->device));
    int rc = ensure_stream(s);
    if (rc) return rc;
    const size_t bytes = (size_t)n * sizeof(int64_t);
    if ((rc = ensure_pinned(s, bytes))) return rc;
    if ((rc = ensure_dev_io(s, bytes))) return rc;
    memcpy(s->pinned, rows_host, bytes);
    CMW_CUDA_OK(cudaMemcpyAsync(s->dev_io, s->pinned, bytes, cudaMemcpyHostToDevice, s->stream));
    return cmw_store_tombstone(h, (const int64_t*)s->dev_io, n, s->stream);
}

const int32_t* cmw_store_kb_gid_dev(const cmw_store* h) {
    return h ? reinterpret_cast<const Store*>(h)->kb_gid : nullptr;
}

int cmw_store_read_rows_f32(cmw_store* h, int64_t row0, int64_t n, float* rows_host, int32_t* kb_gid_host,
                            uint8_t* live_host) {
    CMW_REQUIRE(h != nullptr, "cmw_store_read_rows_f32: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    if (n == 0) return 0;
    CMW_REQUIRE(row0 >= 0 && n > 0 && row0 + n <= s->rows, "cmw_store_read_rows_f32: rows [%lld, %lld) out of range",
                (long long)row0, (long long)(row0 + n));
    std::lock_guard<std::mutex> host_lock(s->host_mu);
    CMW_CUDA_OK(cudaSetDevice(s->device));
    cudaStream_t st;
    int rc = get_stream(s, &st);
    if (rc) return rc;
    if (rows_host != nullptr) {
        CMW_REQUIRE(s->f32 != nullptr, "cmw_store_read_rows_f32: the store keeps no fp32 tiles");
        CMW_CUDA_OK(cudaMemcpyAsync(rows_host, s->f32 + (size_t)row0 * s->dim, (size_t)n * s->dim * sizeof(float),
                                    cudaMemcpyDeviceToHost, st));
    }
    if (kb_gid_host != nullptr)
        CMW_CUDA_OK(cudaMemcpyAsync(kb_gid_host, s->kb_gid + row0, (size_t)n * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, st));
    std::vector<float> live;
    if (live_host != nullptr) {
        live.resize((size_t)n);
        CMW_CUDA_OK(cudaMemcpyAsync(live.data(), s->live + row0, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CMW_CUDA_OK(cudaStreamSynchronize(st));
    if (live_host != nullptr)
        for (int64_t i = 0; i < n; ++i) live_host[i] = (live[i] == live[i]) ? 1 : 0;
    return 0;
}

}  // extern "C"
