// Several GPUs driven by ONE host thread (the reference's caller is a single Python process: runner.py:70-176,
// main.py:299-308): fork/join of per-device product launches on the first shard's stream, partial results combined
// over NVLink peer memory.  No NCCL, no host synchronisation: the caller synchronises shard 0's stream.
#include <mutex>

#include "kmb_common.cuh"

namespace kmb {

int launch_count();          // kmb_api.cu
void set_launch_count(int n);

namespace {

constexpr int MAX_DEVICES = 64;
constexpr int MAX_PARTS = 16;

std::mutex g_multi_mutex;
cudaEvent_t g_fork_event[MAX_DEVICES], g_join_event[MAX_DEVICES];   // lazily created, one pair per device

int device_events(int dev, cudaEvent_t* fork_ev, cudaEvent_t* join_ev) {   // current device must be `dev`
    if (dev < 0 || dev >= MAX_DEVICES) return set_error(KMB_ERR_INVALID, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_multi_mutex);
    if (!g_fork_event[dev]) {
        KMB_CUDA_CHECK(cudaEventCreateWithFlags(&g_fork_event[dev], cudaEventDisableTiming));
        KMB_CUDA_CHECK(cudaEventCreateWithFlags(&g_join_event[dev], cudaEventDisableTiming));
    }
    *fork_ev = g_fork_event[dev];
    *join_ev = g_join_event[dev];
    return KMB_OK;
}

struct PartList {
    const float* p[MAX_PARTS];
};

// out[i] = parts[0][i] + parts[1][i] + ... in a fixed order (deterministic); parts[k] may live on a peer GPU
// (loads travel over NVLink / NVSwitch).  n4 = n / 4 float4 groups + a scalar tail.
__global__ void __launch_bounds__(256) reduce_parts_kernel(float* __restrict__ out, const PartList parts, int n_parts, long long n) {
    const long long n4 = n / 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
        float4 acc = __ldcv(reinterpret_cast<const float4*>(parts.p[0]) + i);
        for (int k = 1; k < n_parts; ++k) {
            const float4 v = __ldcv(reinterpret_cast<const float4*>(parts.p[k]) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
    for (long long i = 4 * n4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        float acc = __ldcv(parts.p[0] + i);
        for (int k = 1; k < n_parts; ++k) acc += __ldcv(parts.p[k] + i);
        out[i] = acc;
    }
}

int check_shards(const kmb_device_shard* shards, int n_shards) {
    if (!shards || n_shards < 1 || n_shards > MAX_PARTS) return set_error(KMB_ERR_INVALID, "need 1..%d device shards (got %d)", MAX_PARTS, n_shards);
    for (int s = 0; s < n_shards; ++s)
        for (int t = 0; t < s; ++t)
            if (shards[s].device == shards[t].device) return set_error(KMB_ERR_INVALID, "device %d listed twice", shards[s].device);
    return KMB_OK;
}

// Everything enqueued on the other shards' streams starts after what shard 0's stream holds now ...
int fork_from_first(const kmb_device_shard* shards, int n_shards) {
    if (n_shards == 1) return KMB_OK;
    cudaEvent_t fork_ev, join_ev;
    KMB_CUDA_CHECK(cudaSetDevice(shards[0].device));
    if (int rc = device_events(shards[0].device, &fork_ev, &join_ev)) return rc;
    KMB_CUDA_CHECK(cudaEventRecord(fork_ev, static_cast<cudaStream_t>(shards[0].stream)));
    for (int s = 1; s < n_shards; ++s) {
        KMB_CUDA_CHECK(cudaSetDevice(shards[s].device));
        KMB_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(shards[s].stream), fork_ev, 0));
    }
    return KMB_OK;
}
// ... and shard 0's stream continues only when all of them are done.
int join_on_first(const kmb_device_shard* shards, int n_shards) {
    for (int s = 1; s < n_shards; ++s) {
        cudaEvent_t fork_ev, join_ev;
        KMB_CUDA_CHECK(cudaSetDevice(shards[s].device));
        if (int rc = device_events(shards[s].device, &fork_ev, &join_ev)) return rc;
        KMB_CUDA_CHECK(cudaEventRecord(join_ev, static_cast<cudaStream_t>(shards[s].stream)));
        KMB_CUDA_CHECK(cudaSetDevice(shards[0].device));
        KMB_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(shards[0].stream), join_ev, 0));
    }
    return KMB_OK;
}

struct DeviceRestore {
    int dev = -1;
    DeviceRestore() { cudaGetDevice(&dev); }
    ~DeviceRestore() { if (dev >= 0) cudaSetDevice(dev); }
};

}  // namespace
}  // namespace kmb

using namespace kmb;

extern "C" {

int kmb_enable_peer_access(const int* devices, int n_devices) {
    if (!devices || n_devices < 1) return set_error(KMB_ERR_INVALID, "no devices");
    DeviceRestore restore;
    for (int a = 0; a < n_devices; ++a) {
        KMB_CUDA_CHECK(cudaSetDevice(devices[a]));
        for (int b = 0; b < n_devices; ++b) {
            if (a == b) continue;
            int can = 0;
            KMB_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
            if (!can) return set_error(KMB_ERR_UNSUPPORTED, "device %d cannot access the memory of device %d (no NVLink / PCIe peer path)", devices[a], devices[b]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); continue; }   // e.g. by the caller's framework
            KMB_CUDA_CHECK(e);
        }
    }
    return KMB_OK;
}

int kmb_reduce_parts_f32(float* out, const float* const* parts, int n_parts, int64_t n, void* stream) {
    if (!out || !parts || n_parts < 1 || n_parts > MAX_PARTS || n < 0) return set_error(KMB_ERR_INVALID, "bad arguments (1..%d parts)", MAX_PARTS);
    if (n == 0) return KMB_OK;
    PartList pl;
    for (int k = 0; k < n_parts; ++k) {
        if (!parts[k] || reinterpret_cast<uintptr_t>(parts[k]) % 16) return set_error(KMB_ERR_INVALID, "part %d is NULL or not 16-byte aligned", k);
        pl.p[k] = parts[k];
    }
    if (reinterpret_cast<uintptr_t>(out) % 16) return set_error(KMB_ERR_INVALID, "out is not 16-byte aligned");
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;
    const long long want = (n / 4 + 255) / 256 + 1;
    const int grid = static_cast<int>(want < 4LL * sms ? want : 4LL * sms);
    reduce_parts_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, pl, n_parts, n);
    KMB_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return KMB_OK;
}

int kmb_product_sym_multi_f32(const kmb_device_shard* shards, int n_shards, float* out, int64_t n, int D, int kernel_id) {
    if (int rc = check_shards(shards, n_shards)) return rc;
    if (!out) return set_error(KMB_ERR_INVALID, "out is NULL");
    DeviceRestore restore;
    if (int rc = fork_from_first(shards, n_shards)) return rc;
    int launches = 0;
    const float* parts[MAX_PARTS];
    for (int s = 0; s < n_shards; ++s) {
        const kmb_device_shard& sh = shards[s];
        KMB_CUDA_CHECK(cudaSetDevice(sh.device));
        float* part = n_shards == 1 ? out : sh.out;
        if (int rc = kmb_product_sym_f32(sh.y, sh.b, part, n, D, kernel_id, s, n_shards, sh.workspace, sh.workspace_bytes, sh.stream)) return rc;
        launches += launch_count();
        parts[s] = part;
    }
    if (n_shards > 1) {
        if (int rc = join_on_first(shards, n_shards)) return rc;
        KMB_CUDA_CHECK(cudaSetDevice(shards[0].device));
        if (int rc = kmb_reduce_parts_f32(out, parts, n_shards, n, shards[0].stream)) return rc;
        launches += 1;
    }
    set_launch_count(launches);
    return KMB_OK;
}

int kmb_product_rows_multi_f32(const kmb_device_shard* shards, int n_shards, int64_t n_sources, int D, int E, int kernel_id,
                               int flags, int path) {
    if (int rc = check_shards(shards, n_shards)) return rc;
    DeviceRestore restore;
    if (int rc = fork_from_first(shards, n_shards)) return rc;
    int launches = 0;
    for (int s = 0; s < n_shards; ++s) {
        const kmb_device_shard& sh = shards[s];
        KMB_CUDA_CHECK(cudaSetDevice(sh.device));
        if (int rc = kmb_product_f32(sh.x, sh.y, sh.b, sh.out, sh.n_targets, n_sources, D, E, kernel_id, flags | sh.flags, path, sh.row_offset,
                                     sh.workspace, sh.workspace_bytes, sh.stream)) return rc;
        launches += launch_count();
    }
    if (int rc = join_on_first(shards, n_shards)) return rc;
    set_launch_count(launches);
    return KMB_OK;
}

}  // extern "C"
