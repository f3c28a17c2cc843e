// C ABI (include/bhgeo.h) over the sm_100a trace kernels.  No torch, no Python: plain pointers and sizes,
// so the library loads with ctypes.CDLL from any CPython (Blender 4.1's bundled one included,
// bl_info at /root/reference/raytracer/RelativisticRenderEngine.py:19).
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/bhgeo.h"
#include "trace_kernel.cuh"
#include "aux_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define BHG_CUDA(call)                                                                                 \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? BHG_ERR_OUT_OF_MEMORY                      \
                        : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ||             \
                           e_ == cudaErrorInvalidDevice)                                               \
                            ? BHG_ERR_NO_DEVICE                                                        \
                            : BHG_ERR_CUDA;                                                            \
            return fail(code_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                              \
    } while (0)

// Restores the calling thread's current CUDA device on scope exit: the library selects `device` for its own
// launches but must not change what the host application (torch, Blender) believes is current.
struct DeviceRestore {
    int prev = -1;
    DeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
};

// Small persistent pool for host-side copies between pageable user arrays and pinned bounce buffers: one memcpy
// thread reaches ~10 GB/s, PCIe needs > 50.
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool;  // intentionally never destroyed: its threads live until process exit
        return *p;
    }
    void copy(void* dst, const void* src, size_t bytes) {
        if (bytes < (1u << 20) || workers_.empty()) {
            memcpy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> one_copy_at_a_time(call_mu_);  // host threads of different devices share the pool
        std::unique_lock<std::mutex> lk(mu_);
        const size_t parts = workers_.size() + 1;
        const size_t slice = ((bytes / parts) + 4095) & ~size_t(4095);
        dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes; slice_ = slice;
        pending_ = (int)workers_.size();
        ++generation_;
        cv_.notify_all();
        lk.unlock();
        run_slice(parts - 1);  // the caller takes the last slice
        lk.lock();
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    CopyPool() {
        unsigned hc = std::thread::hardware_concurrency();
        int n = (int)(hc / 2);
        if (n > 7) n = 7;
        if (n < 1) n = 1;
        for (int i = 0; i < n; i++) workers_.emplace_back([this, i] { loop(i); });
        for (auto& t : workers_) t.detach();  // process-lifetime helpers
    }
    void run_slice(size_t i) {
        const size_t b = i * slice_;
        if (b >= bytes_) return;
        const size_t m = (bytes_ - b < slice_) ? bytes_ - b : slice_;
        memcpy(dst_ + b, src_ + b, m);
    }
    void loop(int i) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return generation_ != seen; });
            seen = generation_;
            lk.unlock();
            run_slice((size_t)i);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0;
    int pending_ = 0;
    unsigned long long generation_ = 0;
};

bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

constexpr int kMaxDevices = 64;
constexpr int kQueueSlots = 256;

struct DeviceCtx {
    std::mutex mu;
    bool ready = false;
    int sm_count = 0;
    int blocks_per_sm[2][3] = {{0, 0, 0}, {0, 0, 0}};  // [mode][layout]
    unsigned long long* queue_slots = nullptr;   // kQueueSlots work-queue heads
    std::atomic<unsigned> next_slot{0};
    // staging buffers of the host entry point (grow-only)
    std::mutex host_mu;
    void* stage = nullptr;
    size_t stage_bytes = 0;
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    cudaStream_t courier_stream = nullptr;   // bhg_trace_frame_shard_f64 (created on first use)
    bool courier_loaded = false;             // its kernels are resident (lazy module loading)
    // pinned bounce buffers for pageable user arrays (3 pipeline slots), grow-only
    void* bounce = nullptr;
    size_t bounce_bytes = 0;
    cudaEvent_t slot_in_done[3] = {nullptr, nullptr, nullptr}, slot_out_done[3] = {nullptr, nullptr, nullptr};
    long long* totals = nullptr;  // 3 int64 for bhg_sum_counters
};

DeviceCtx g_ctx[kMaxDevices];

template <int NK, int IN>
int query_occupancy(int* out) {
    BHG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, bhg::trace_kernel<NK, IN, false, false, true>, BHG_BLOCK, 0));
    return 0;
}

int ensure_device(int device, DeviceCtx** out) {
    if (device < 0 || device >= kMaxDevices) return fail(BHG_ERR_INVALID_ARGUMENT, "device ordinal %d out of range", device);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(BHG_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device >= count) return fail(BHG_ERR_NO_DEVICE, "device %d requested but only %d present", device, count);
    BHG_CUDA(cudaSetDevice(device));
    DeviceCtx& c = g_ctx[device];
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) {
        cudaDeviceProp prop;
        BHG_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(BHG_ERR_NO_DEVICE, "device %d is sm_%d%d; this build contains sm_100a code only", device,
                        prop.major, prop.minor);
        c.sm_count = prop.multiProcessorCount;
        int rc;
        if ((rc = query_occupancy<4, bhg::IN_SOA>(&c.blocks_per_sm[0][0]))) return rc;
        if ((rc = query_occupancy<4, bhg::IN_AOS>(&c.blocks_per_sm[0][1]))) return rc;
        if ((rc = query_occupancy<4, bhg::IN_AOS_F32>(&c.blocks_per_sm[0][2]))) return rc;
        if ((rc = query_occupancy<3, bhg::IN_AOS_F32>(&c.blocks_per_sm[1][2]))) return rc;
        if ((rc = query_occupancy<3, bhg::IN_SOA>(&c.blocks_per_sm[1][0]))) return rc;
        if ((rc = query_occupancy<3, bhg::IN_AOS>(&c.blocks_per_sm[1][1]))) return rc;
        BHG_CUDA(cudaMalloc(&c.queue_slots, kQueueSlots * sizeof(unsigned long long)));
        BHG_CUDA(cudaMalloc(&c.totals, 3 * sizeof(long long)));
        // keep stream-ordered scratch (camera ray buffers) in the pool instead of returning it to the OS at
        // every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        c.ready = true;
    }
    *out = &c;
    return 0;
}

int validate(const bhg_params* p, long long n) {
    if (!p) return fail(BHG_ERR_INVALID_ARGUMENT, "params is NULL");
    if (n < 0) return fail(BHG_ERR_INVALID_ARGUMENT, "n = %lld is negative", n);
    if (n > 2147483647LL) return fail(BHG_ERR_INVALID_ARGUMENT, "n = %lld exceeds 2^31-1 rays per call", n);
    if (!(p->M > 0.0) || !std::isfinite(p->M)) return fail(BHG_ERR_INVALID_ARGUMENT, "M must be finite and > 0");
    if (!(p->r_sphere > 2.0 * p->M)) return fail(BHG_ERR_INVALID_ARGUMENT, "r_sphere must exceed r_s = 2 M");
    if (!(p->rtol > 0.0) || !(p->atol >= 0.0)) return fail(BHG_ERR_INVALID_ARGUMENT, "rtol must be > 0 and atol >= 0");
    if (!(p->max_step > 0.0)) return fail(BHG_ERR_INVALID_ARGUMENT, "max_step must be > 0 (use +inf for unbounded)");
    if (!(p->eps_horizon >= 0.0)) return fail(BHG_ERR_INVALID_ARGUMENT, "eps_horizon must be >= 0");
    if (p->mode != BHG_MODE_PARITY && p->mode != BHG_MODE_PLANE)
        return fail(BHG_ERR_INVALID_ARGUMENT, "unknown mode %d", p->mode);
    if (p->refill_threshold < 0 || p->refill_threshold > 32)
        return fail(BHG_ERR_INVALID_ARGUMENT, "refill_threshold must be in 0..32");
    if (p->image_width < 0) return fail(BHG_ERR_INVALID_ARGUMENT, "image_width must be >= 0");
    if (p->coords != BHG_COORDS_SCHWARZSCHILD && p->coords != BHG_COORDS_ISOTROPIC)
        return fail(BHG_ERR_INVALID_ARGUMENT, "unknown coords %d", p->coords);
    double lam = p->lambda_max;
    if (!(lam > 0.0) && !std::isfinite(p->r_sphere))
        return fail(BHG_ERR_INVALID_ARGUMENT, "lambda_max must be given when r_sphere is infinite");
    return 0;
}

int convert_camera(const bhg_camera* cam, double r_sphere, bhg::Camera* out) {
    if (!cam) return fail(BHG_ERR_INVALID_ARGUMENT, "camera is NULL");
    if (cam->width <= 0 || cam->height <= 0) return fail(BHG_ERR_INVALID_ARGUMENT, "camera width/height must be > 0");
    if (cam->first_ray < 0 || cam->reserved != 0 || (cam->jitter != 0 && cam->jitter != 1))
        return fail(BHG_ERR_INVALID_ARGUMENT, "camera first_ray must be >= 0, jitter 0 or 1, reserved 0");
    if (!std::isfinite(r_sphere)) return fail(BHG_ERR_INVALID_ARGUMENT, "camera entry needs a finite r_sphere");
    for (int i = 0; i < 3; i++) out->origin[i] = cam->origin[i];
    for (int i = 0; i < 9; i++) out->rot[i] = cam->rotation[i];
    out->fov_x = cam->fov_x; out->fov_y = cam->fov_y;
    out->r_sphere = r_sphere;
    out->first_ray = cam->first_ray;
    out->seed = cam->seed;
    out->width = cam->width; out->height = cam->height;
    out->jitter = cam->jitter;
    return 0;
}

// The pre-pass (prepare_kernel) is the default; BHG_PREP=0 keeps the initialisation inside the trace kernel (A/B runs)
// Cost binning of bundles that carry no image order (cost_key_kernel): on unless BHG_BIN=0; BHG_BIN=2 forces it even
// when the caller gave the image_width hint (experiments)
int bin_mode() {
    static const int v = [] {
        const char* e = getenv("BHG_BIN");
        return e ? atoi(e) : 1;
    }();
    return v;
}

template <int IN>
void launch_cost_passes(int sample_blocks, int blocks, cudaStream_t stream, const bhg::TraceArgs& a, unsigned char* keys,
                        int* hist, int32_t* sorted, int force) {
    bhg::cost_sample_kernel<IN><<<sample_blocks, 256, 0, stream>>>(a, hist);
    bhg::cost_decide_kernel<<<1, 32, 0, stream>>>(hist, force, a.idle_budget, a.idle_budget < 32 ? a.idle_budget : 32);
    bhg::cost_key_kernel<IN><<<blocks, 256, 0, stream>>>(a, keys, hist);
    bhg::cost_offsets_kernel<<<1, 32, 0, stream>>>(hist);
    bhg::cost_scatter_kernel<<<blocks, 256, 0, stream>>>(a.n, a.order, keys, hist, sorted);
}

// long rays first (trace_kernel.cuh TraceArgs::hot_list): on unless BHG_HOT=0
bool hot_enabled() {
    static const int v = [] {
        const char* e = getenv("BHG_HOT");
        return (e && atoi(e) == 0) ? 0 : 1;
    }();
    return v != 0;
}

bool prep_enabled() {
    static const int v = [] {
        const char* e = getenv("BHG_PREP");
        return (e && atoi(e) == 0) ? 0 : 1;
    }();
    return v != 0;
}

template <int NK, int IN, bool DISK, bool POLY>
void launch_trace_variant(bool prep, bool staged, int blocks, cudaStream_t stream, const bhg::TraceArgs& a) {
    if constexpr (IN == bhg::IN_AOS && !DISK && !POLY) {
        if (prep && staged) {  // exit states go to another GPU's memory: coalesce them through the staging tiles
            bhg::trace_kernel<NK, IN, false, false, true, true><<<blocks, BHG_BLOCK, 0, stream>>>(a);
            return;
        }
    }
    if (prep) bhg::trace_kernel<NK, IN, DISK, POLY, true><<<blocks, BHG_BLOCK, 0, stream>>>(a);
    else bhg::trace_kernel<NK, IN, DISK, POLY, false><<<blocks, BHG_BLOCK, 0, stream>>>(a);
}

// true when `p` is device memory of ANOTHER GPU (a bhg_ipc_open / peer mapping)
bool is_remote(const void* p, int device) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice && at.device != device;
}

template <int NK>
void launch_prepare_variant(int in_kind, bool from_camera, int blocks, cudaStream_t stream, const bhg::TraceArgs& a,
                            const bhg::Camera& cam, double* ray_pos, double* ray_dir) {
    if (from_camera) bhg::prepare_kernel<NK, bhg::IN_CAMERA><<<blocks, 256, 0, stream>>>(a, cam, ray_pos, ray_dir);
    else if (in_kind == bhg::IN_AOS) bhg::prepare_kernel<NK, bhg::IN_AOS><<<blocks, 256, 0, stream>>>(a, cam, nullptr, nullptr);
    else if (in_kind == bhg::IN_AOS_F32) bhg::prepare_kernel<NK, bhg::IN_AOS_F32><<<blocks, 256, 0, stream>>>(a, cam, nullptr, nullptr);
    else bhg::prepare_kernel<NK, bhg::IN_SOA><<<blocks, 256, 0, stream>>>(a, cam, nullptr, nullptr);
}

// sharded frames (bhg_trace_frame_shard_f64): band progress counters for the courier and SMs left free for it
struct ShardHook {
    int* band_done;
    long long band_rays;
    int reserve_sms;
    long long map_first, map_stride;   // input mapping (TraceArgs::map_*): bands of band_rays, stride 0 = compact input
};

// in_kind: bhg::IN_SOA / IN_AOS / IN_AOS_F32.  `cam` != NULL: the rays come from the camera description (in / in_dir
// are ignored; outputs are AoS) - parity mode needs no ray buffer at all, plane mode keeps one for its exit frame.
int launch_trace(DeviceCtx& c, const double* in, const double* in_dir, double* out, double* out_dir, int32_t* status,
                 int32_t* counters, const int32_t* order, long long n, int in_kind, int image_width,
                 const bhg_params* p, cudaStream_t stream, const bhg_extras* ex = nullptr,
                 const bhg::Camera* cam = nullptr, const ShardHook* hook = nullptr) {
    if (n == 0) return 0;
    const bool disk = ex && ex->disk_xy && ex->disk_r_out > 0.0;
    const bool poly = ex && ex->poly_n >= 2 && ex->poly_xyz && ex->poly_count;
    if (poly && (p->mode != BHG_MODE_PARITY || in_kind != bhg::IN_AOS || !(p->lambda_max > 0.0)))
        return fail(BHG_ERR_INVALID_ARGUMENT, "the polyline output needs parity mode, the float64 AOS layout and an explicit lambda_max");
    if (disk && p->mode != BHG_MODE_PARITY)
        return fail(BHG_ERR_INVALID_ARGUMENT, "the disk-crossing event is available in parity mode only");
    if (disk && in_kind != bhg::IN_AOS)
        return fail(BHG_ERR_INVALID_ARGUMENT, "the disk-crossing event needs the float64 AOS layout");
    if (disk && !(ex->disk_r_in >= 0.0 && ex->disk_r_in <= ex->disk_r_out))
        return fail(BHG_ERR_INVALID_ARGUMENT, "the disk annulus needs 0 <= disk_r_in <= disk_r_out");
    bhg::TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.in_dir = in_dir; a.out = out; a.out_dir = out_dir;
    if (hook) {
        a.band_done = hook->band_done;
        a.band_rays = hook->band_rays;
        if (hook->map_stride > 0) { a.map_band = hook->band_rays; a.map_first = hook->map_first; a.map_stride = hook->map_stride; }
    }
    a.status = status; a.counters = counters; a.order = order;
    a.n = n;
    a.rs = 2.0 * p->M;
    a.r_hor = a.rs + p->eps_horizon;
    a.r_sphere = p->r_sphere;
    a.coords = p->coords;
    // isotropic boundary: radii the caller states (sphere of influence, disk annulus) are isotropic radii; the
    // integration runs on the Schwarzschild radius r = rho (1 + r_s / 4 rho)^2
    auto schw_radius = [&](double rho) {
        if (!(p->coords == BHG_COORDS_ISOTROPIC) || !std::isfinite(rho) || !(rho > 0.0)) return rho;
        const double q = 1.0 + a.rs / (4.0 * rho);
        return rho * q * q;
    };
    a.r_sphere = schw_radius(a.r_sphere);
    a.has_outer = std::isfinite(p->r_sphere) ? 1 : 0;
    a.rtol = p->rtol; a.atol = p->atol; a.max_step = p->max_step;
    a.has_max_step = std::isfinite(a.max_step) ? 1 : 0;
    a.atol_over_rtol = a.atol / a.rtol; a.inv_rtol2 = 1.0 / (a.rtol * a.rtol);
    a.lambda_max = p->lambda_max > 0.0 ? p->lambda_max : 10.0 * p->r_sphere;
    // refill policy: explicit lane threshold, else the adaptive idle budget (lane-iterations per service)
    a.refill_threshold = p->refill_threshold;
    a.idle_budget = 96;  // measured optimum across coherent and incoherent workloads (profiles/r1i_refill_policy.txt)
    if (const char* e = getenv("BHG_IDLE_BUDGET")) {
        int v = atoi(e);
        if (v > 0) a.idle_budget = v;
    }
    a.tile_width = (image_width > 0 && image_width % 4 == 0 && n % (8LL * image_width) == 0 && !order) ? image_width : 0;
    if (a.tile_width > 0) {
        // band = slot / d, d = 8 tile_width, for every slot < 2^31: (slot * ceil(2^(31+l) / d)) >> (31 + l), l = ceil(log2 d)
        const unsigned long long d = 8ULL * (unsigned long long)a.tile_width;
        int l = 0;
        while ((1ULL << l) < d) l++;
        a.tile_shift = 31 + l;
        a.tile_magic = (unsigned long long)((((unsigned __int128)1 << a.tile_shift) + d - 1) / d);
    }
    if (disk) { a.disk_r_in = schw_radius(ex->disk_r_in); a.disk_r_out = schw_radius(ex->disk_r_out); a.disk_xy = ex->disk_xy; }
    if (poly) {
        a.poly_n = ex->poly_n;
        a.poly_dt = a.lambda_max / (ex->poly_n - 1);
        a.poly_xyz = ex->poly_xyz;
        a.poly_count = ex->poly_count;
    }
    // work-queue head: with the pre-pass (the default) it lives in this launch's own scratch, set below; the ring of
    // kQueueSlots heads serves BHG_PREP=0 only (launches further apart than the ring never overlap on one stream)
    unsigned slot = c.next_slot.fetch_add(1) % kQueueSlots;
    a.queue_head = c.queue_slots + slot;
    BHG_CUDA(cudaMemsetAsync(a.queue_head, 0, sizeof(unsigned long long), stream));
    const int mode = p->mode;
    // ---- cost binning: a bundle without image order is served costliest class first (unless found coherent)
    void* bin_scratch = nullptr;
    const int bm = bin_mode();
    if (!cam && n >= 4096 && n < (1LL << 31) && ((bm == 1 && a.tile_width == 0) || bm == 2)) {
        const size_t keys_bytes = ((size_t)n + 255) & ~(size_t)255;
        BHG_CUDA(cudaMallocAsync(&bin_scratch, keys_bytes + 512 + (size_t)n * sizeof(int32_t), stream));
        unsigned char* keys = (unsigned char*)bin_scratch;
        int* hist = (int*)(keys + keys_bytes);          // counts, cursors, sample statistics, keep flag (trace_kernel.cuh)
        int32_t* sorted = (int32_t*)(keys + keys_bytes + 512);
        BHG_CUDA(cudaMemsetAsync(hist, 0, 512, stream));
        long long kb = (n + 255) / 256;
        if (kb > c.sm_count * 8LL) kb = c.sm_count * 8LL;
        long long sb = (n / (32 * bhg::COST_SAMPLE_STRIDE) + 7) / 8;
        if (sb < 1) sb = 1;
        if (sb > c.sm_count * 8LL) sb = c.sm_count * 8LL;
        bhg::TraceArgs ka = a;
        if (bm == 2) ka.tile_width = 0;
        const int force = bm == 2 ? 1 : 0;
        if (in_kind == bhg::IN_AOS) launch_cost_passes<bhg::IN_AOS>((int)sb, (int)kb, stream, ka, keys, hist, sorted, force);
        else if (in_kind == bhg::IN_AOS_F32) launch_cost_passes<bhg::IN_AOS_F32>((int)sb, (int)kb, stream, ka, keys, hist, sorted, force);
        else launch_cost_passes<bhg::IN_SOA>((int)sb, (int)kb, stream, ka, keys, hist, sorted, force);
        g_launches.fetch_add(5);
        BHG_CUDA(cudaGetLastError());
        a.sorted = sorted;
        a.sort_keep = hist + 2 * bhg::COST_BINS + 2;
        if (a.refill_threshold == 0 && !getenv("BHG_IDLE_BUDGET")) a.idle_budget_dev = hist + 2 * bhg::COST_BINS + 3;
        if (bm == 2) a.tile_width = 0;
    }
    // ---- pre-pass: prepared rays in queue order (stream-ordered scratch)
    const bool prep = prep_enabled() || cam != nullptr;
    double* cam_rays = nullptr;
    if (prep) {
        const int planes = (mode == BHG_MODE_PARITY) ? 6 : 4;
        // prepared records + the long-ray list: [count 256 B][mask][list]
        const size_t rec_bytes = (size_t)n * planes * sizeof(double2);
        const size_t mask_bytes = ((((size_t)n + 31) / 32) * 4 + 255) & ~(size_t)255;
        // the tail of a launch is a fixed ~0.05 ms: worth the list (+0.6 % pre-pass work) only below ~30 rays per lane
        const bool hot = hot_enabled() && n <= (1LL << 21);
        BHG_CUDA(cudaMallocAsync((void**)&a.prep, rec_bytes + 256 + (hot ? mask_bytes + (size_t)n * 4 : 0), stream));
        // header of the scratch: [0] this launch's OWN work-queue head (launches on different streams, or a captured
        // graph replayed next to eager launches, share nothing), [1] the long-ray count
        char* hb = (char*)a.prep + rec_bytes;
        a.queue_head = (unsigned long long*)hb;
        if (hot) {
            a.hot_count = (unsigned long long*)hb + 1;
            a.hot_mask = (unsigned int*)(hb + 256);
            a.hot_list = (int32_t*)(hb + 256 + mask_bytes);
        }
        BHG_CUDA(cudaMemsetAsync(hb, 0, 256 + (hot ? mask_bytes : 0), stream));
        if (cam && mode == BHG_MODE_PLANE) {  // the orbital-plane frame is rebuilt from the flat entry state at the exit
            BHG_CUDA(cudaMallocAsync((void**)&cam_rays, (size_t)n * 48, stream));
            a.in = cam_rays;
            a.in_dir = cam_rays + 3 * n;
        }
        long long pb = (n + 255) / 256;
        if (pb > c.sm_count * 16LL) pb = c.sm_count * 16LL;
        bhg::Camera none;
        memset(&none, 0, sizeof(none));
        const bhg::Camera& cc = cam ? *cam : none;
        if (mode == BHG_MODE_PARITY) launch_prepare_variant<4>(in_kind, cam != nullptr, (int)pb, stream, a, cc, nullptr, nullptr);
        else launch_prepare_variant<3>(in_kind, cam != nullptr, (int)pb, stream, a, cc, cam_rays, cam_rays ? cam_rays + 3 * n : nullptr);
        g_launches.fetch_add(1);
        BHG_CUDA(cudaGetLastError());
    }
    long long want_blocks = (n + BHG_BLOCK - 1) / BHG_BLOCK;
    long long max_blocks = (long long)c.sm_count * c.blocks_per_sm[mode][in_kind];
    if (const char* e = getenv("BHG_BLOCKS_PER_SM")) {  // tuning experiments: fewer resident warps
        int v = atoi(e);
        if (v > 0 && v < c.blocks_per_sm[mode][in_kind]) max_blocks = (long long)c.sm_count * v;
    }
    if (hook) {
        const long long keep = (long long)(c.sm_count - hook->reserve_sms) * c.blocks_per_sm[mode][in_kind];
        if (keep >= 1 && keep < max_blocks) max_blocks = keep;
    }
    int blocks = (int)(want_blocks < max_blocks ? want_blocks : max_blocks);
    const bool staged = is_remote(out_dir, (int)(&c - g_ctx));
    if (poly) {
        if (disk) launch_trace_variant<4, bhg::IN_AOS, true, true>(prep, staged, blocks, stream, a);
        else launch_trace_variant<4, bhg::IN_AOS, false, true>(prep, staged, blocks, stream, a);
    } else if (disk) {
        launch_trace_variant<4, bhg::IN_AOS, true, false>(prep, staged, blocks, stream, a);
    } else if (mode == BHG_MODE_PARITY) {
        if (in_kind == bhg::IN_AOS) launch_trace_variant<4, bhg::IN_AOS, false, false>(prep, staged, blocks, stream, a);
        else if (in_kind == bhg::IN_AOS_F32) launch_trace_variant<4, bhg::IN_AOS_F32, false, false>(prep, staged, blocks, stream, a);
        else launch_trace_variant<4, bhg::IN_SOA, false, false>(prep, staged, blocks, stream, a);
    } else {
        if (in_kind == bhg::IN_AOS) launch_trace_variant<3, bhg::IN_AOS, false, false>(prep, staged, blocks, stream, a);
        else if (in_kind == bhg::IN_AOS_F32) launch_trace_variant<3, bhg::IN_AOS_F32, false, false>(prep, staged, blocks, stream, a);
        else launch_trace_variant<3, bhg::IN_SOA, false, false>(prep, staged, blocks, stream, a);
    }
    g_launches.fetch_add(1);
    BHG_CUDA(cudaGetLastError());
    if (bin_scratch) cudaFreeAsync(bin_scratch, stream);
    if (a.prep) cudaFreeAsync(a.prep, stream);
    if (cam_rays) cudaFreeAsync(cam_rays, stream);
    return 0;
}

// primary rays of `cam` into AoS device buffers (streaming kernel, HBM-bound: 48 B/ray written)
int launch_generate(DeviceCtx& c, const bhg::Camera& cam, long long n, double* pos, double* dir, int32_t* hit,
                    cudaStream_t stream) {
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > c.sm_count * 16LL) blocks = c.sm_count * 16LL;
    bhg::generate_rays_kernel<<<(int)blocks, 256, 0, stream>>>(cam, n, pos, dir, hit);
    g_launches.fetch_add(1);
    BHG_CUDA(cudaGetLastError());
    return 0;
}

// the tile hint of a camera call: valid only when the call starts on an 8-row band boundary
int camera_image_width(const bhg::Camera& cam) { return (cam.first_ray % (8LL * cam.width) == 0) ? cam.width : 0; }

// rays per pipeline chunk of the host entry points; BHG_CHUNK_RAYS overrides (tuning)
// (measured on B200, profiles/r1g_chunks.txt: 512 Ki rays is best when rays also travel H2D, 256 Ki when only
// results travel D2H)
long long pick_chunk(long long n, long long band, long long big = 1 << 18) {
    long long chunk = n <= (1 << 16) ? n : (n <= (1 << 20) ? (n + 3) / 4 : big);
    if (const char* e = getenv("BHG_CHUNK_RAYS")) {
        long long v = atoll(e);
        if (v > 0) chunk = v;
    }
    if (chunk > n) chunk = n;
    if (band > 0 && n > chunk) chunk = ((chunk + band - 1) / band) * band;  // whole 8-row bands keep the tile hint
    return chunk;
}

int ensure_stage(DeviceCtx* c, size_t need) {
    if (c->stage_bytes < need) {
        if (c->stage) cudaFree(c->stage);
        c->stage = nullptr;
        c->stage_bytes = 0;
        BHG_CUDA(cudaMalloc(&c->stage, need));
        c->stage_bytes = need;
    }
    for (auto& s : c->streams)
        if (!s) BHG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return 0;
}

// The host entry points queue asynchronous copies into the caller's arrays; on an early error return those must
// not still be in flight when the caller gets its buffers back.
struct DrainStreamsOnExit {
    DeviceCtx* c;
    ~DrainStreamsOnExit() {
        for (auto& s : c->streams)
            if (s) cudaStreamSynchronize(s);
    }
};

}  // namespace

extern "C" {

void bhg_default_params(bhg_params* p) {
    if (!p) return;
    p->M = 1.0;
    p->r_sphere = 60.0;
    p->rtol = 1e-3;
    p->atol = 1e-6;
    p->max_step = std::numeric_limits<double>::infinity();
    p->eps_horizon = 0.01;
    p->lambda_max = 0.0;
    p->mode = BHG_MODE_PARITY;
    p->refill_threshold = 0;
    p->image_width = 0;
    p->coords = BHG_COORDS_SCHWARZSCHILD;
}

int bhg_trace_schwarzschild_f64(const double* in, const double* in_dir, double* out, double* out_dir,
                                int32_t* status, int32_t* counters, const int32_t* order, int64_t n,
                                int32_t layout, const bhg_params* params, int32_t device, void* stream) {
    return bhg_trace_schwarzschild_f64_ex(in, in_dir, out, out_dir, status, counters, order, n, layout, params, nullptr,
                                          device, stream);
}

int bhg_trace_schwarzschild_f64_ex(const double* in, const double* in_dir, double* out, double* out_dir,
                                   int32_t* status, int32_t* counters, const int32_t* order, int64_t n,
                                   int32_t layout, const bhg_params* params, const bhg_extras* extras,
                                   int32_t device, void* stream) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    if (layout != BHG_LAYOUT_SOA && layout != BHG_LAYOUT_AOS) return fail(BHG_ERR_INVALID_ARGUMENT, "unknown layout %d", layout);
    if (n > 0 && (!in || !out || !status)) return fail(BHG_ERR_INVALID_ARGUMENT, "NULL ray buffer");
    if (n > 0 && layout == BHG_LAYOUT_AOS && (!in_dir || !out_dir))
        return fail(BHG_ERR_INVALID_ARGUMENT, "AOS layout needs in_dir and out_dir");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    return launch_trace(*c, in, in_dir, out, out_dir, status, counters, order, n,
                        layout == BHG_LAYOUT_AOS ? bhg::IN_AOS : bhg::IN_SOA, params->image_width, params,
                        (cudaStream_t)stream, extras);
}

int bhg_trace_schwarzschild_f64_host(const double* entry_pos, const double* entry_dir, double* exit_pos,
                                     double* exit_dir, int32_t* status, int32_t* counters, int64_t n,
                                     const bhg_params* params, int32_t device) {
    return bhg_trace_schwarzschild_f64_host_ex(entry_pos, entry_dir, exit_pos, exit_dir, status, counters, n, params,
                                               nullptr, device);
}

int bhg_trace_schwarzschild_f64_host_ex(const double* entry_pos, const double* entry_dir, double* exit_pos,
                                        double* exit_dir, int32_t* status, int32_t* counters, int64_t n,
                                        const bhg_params* params, const bhg_extras* extras, int32_t device) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    if (n > 0 && (!entry_pos || !entry_dir || !exit_pos || !exit_dir || !status))
        return fail(BHG_ERR_INVALID_ARGUMENT, "NULL ray buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(c->host_mu);
    // device staging: pos_in | dir_in | pos_out | dir_out | status | counters
    const size_t vec = (size_t)n * 3 * sizeof(double);
    const bool disk = extras && extras->disk_xy && extras->disk_r_out > 0.0;
    if (extras && extras->poly_n >= 2 && extras->poly_xyz && extras->poly_count) {
        // polyline requests (small batches by nature: n x poly_n x 24 bytes come back) take a simple unpipelined path
        cudaStream_t s0 = nullptr;
        for (auto& st : c->streams)
            if (!st) BHG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        s0 = c->streams[0];
        const size_t pbytes = (size_t)n * extras->poly_n * 24;
        char* dbuf = nullptr;
        const size_t tot = 4 * vec + (size_t)n * 16 + (size_t)n * 4 * 4 + pbytes + 256;
        BHG_CUDA(cudaMallocAsync((void**)&dbuf, tot, s0));
        double* q_pin = (double*)dbuf;
        double* q_din = q_pin + 3 * n;
        double* q_pout = q_din + 3 * n;
        double* q_dout = q_pout + 3 * n;
        double* q_disk = q_dout + 3 * n;
        double* q_poly = q_disk + 2 * n;
        int32_t* q_status = (int32_t*)(q_poly + (size_t)n * extras->poly_n * 3);
        int32_t* q_cnt = q_status + n;      // 2 n
        int32_t* q_pcount = q_cnt + 2 * n;  // n
        bhg_extras ex = *extras;
        ex.disk_xy = disk ? q_disk : nullptr;
        ex.poly_xyz = q_poly;
        ex.poly_count = q_pcount;
        rc = 0;
        auto cp = [&](void* d, const void* sp, size_t b, cudaMemcpyKind kd) {
            if (!rc && cudaMemcpyAsync(d, sp, b, kd, s0) != cudaSuccess) rc = fail(BHG_ERR_CUDA, "polyline staging copy failed");
        };
        cp(q_pin, entry_pos, vec, cudaMemcpyHostToDevice);
        cp(q_din, entry_dir, vec, cudaMemcpyHostToDevice);
        // samples beyond a ray's count are not written by the kernel: all-ones bytes = NaN
        if (!rc && cudaMemsetAsync(q_poly, 0xFF, pbytes, s0) != cudaSuccess) rc = fail(BHG_ERR_CUDA, "polyline memset failed");
        if (!rc) rc = launch_trace(*c, q_pin, q_din, q_pout, q_dout, q_status, counters ? q_cnt : nullptr, nullptr, n,
                                   bhg::IN_AOS, params->image_width, params, s0, &ex);
        cp(exit_pos, q_pout, vec, cudaMemcpyDeviceToHost);
        cp(exit_dir, q_dout, vec, cudaMemcpyDeviceToHost);
        cp(status, q_status, (size_t)n * 4, cudaMemcpyDeviceToHost);
        if (counters) cp(counters, q_cnt, (size_t)n * 8, cudaMemcpyDeviceToHost);
        if (disk) cp(extras->disk_xy, q_disk, (size_t)n * 16, cudaMemcpyDeviceToHost);
        cp(extras->poly_xyz, q_poly, pbytes, cudaMemcpyDeviceToHost);
        cp(extras->poly_count, q_pcount, (size_t)n * 4, cudaMemcpyDeviceToHost);
        cudaFreeAsync(dbuf, s0);
        if (cudaStreamSynchronize(s0) != cudaSuccess && !rc) rc = fail(BHG_ERR_CUDA, "polyline trace failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    const size_t need = 4 * vec + (size_t)n * 3 * sizeof(int32_t) + (disk ? (size_t)n * 16 : 0) + 1024;
    if ((rc = ensure_stage(c, need))) return rc;
    DrainStreamsOnExit drain_streams_on_exit{c};
    char* base = (char*)c->stage;
    double* d_disk = (double*)(base + 4 * vec + (((size_t)n * 3 * sizeof(int32_t) + 15) / 16) * 16);
    double* d_pin = (double*)base;
    double* d_din = (double*)(base + vec);
    double* d_pout = (double*)(base + 2 * vec);
    double* d_dout = (double*)(base + 3 * vec);
    int32_t* d_status = (int32_t*)(base + 4 * vec);
    int32_t* d_cnt = d_status + n;  // 2 n
    // chunked pipeline over 3 streams: H2D(i+1) overlaps trace(i) overlaps D2H(i-1)
    // Pageable user arrays (plain numpy) are staged through pinned bounce slots with a multi-threaded memcpy;
    // pinned arrays (bhg_host_alloc) are copied directly.
    const bool bounce = is_pageable(entry_pos) || is_pageable(entry_dir) || is_pageable(exit_pos) ||
                        is_pageable(exit_dir) || is_pageable(status) || is_pageable(counters) ||
                        (disk && is_pageable(extras->disk_xy));
    const long long chunk = pick_chunk(n, params->image_width > 0 ? 8LL * params->image_width : 0, bounce ? 1 << 18 : 1 << 19);
    // per-slot bounce layout: pos_in | dir_in | pos_out | dir_out | disk | status | counters(2)
    const size_t slot_bytes = (size_t)chunk * (24 * 4 + 16 + 4 * 3) + 256;
    char* bb = nullptr;
    if (bounce) {
        if (c->bounce_bytes < 3 * slot_bytes) {
            if (c->bounce) cudaFreeHost(c->bounce);
            c->bounce = nullptr;
            c->bounce_bytes = 0;
            BHG_CUDA(cudaHostAlloc(&c->bounce, 3 * slot_bytes, cudaHostAllocPortable));
            c->bounce_bytes = 3 * slot_bytes;
        }
        for (int i = 0; i < 3; i++) {
            if (!c->slot_in_done[i]) BHG_CUDA(cudaEventCreateWithFlags(&c->slot_in_done[i], cudaEventDisableTiming));
            if (!c->slot_out_done[i]) BHG_CUDA(cudaEventCreateWithFlags(&c->slot_out_done[i], cudaEventDisableTiming));
        }
        bb = (char*)c->bounce;
    }
    CopyPool& pool = CopyPool::get();
    auto slot_ptrs = [&](int slot, double*& pin, double*& din, double*& pout, double*& dout, double*& dsk,
                         int32_t*& st, int32_t*& cn) {
        char* sb = bb + (size_t)slot * slot_bytes;
        pin = (double*)sb;
        din = pin + 3 * chunk;
        pout = din + 3 * chunk;
        dout = pout + 3 * chunk;
        dsk = dout + 3 * chunk;
        st = (int32_t*)(dsk + 2 * chunk);
        cn = st + chunk;
    };
    // copies a finished chunk from its bounce slot to the user's arrays
    auto drain = [&](long long b, int slot) -> int {
        const long long m = (n - b < chunk) ? (n - b) : chunk;
        double *pin, *din, *pout, *dout, *dsk;
        int32_t *st, *cn;
        slot_ptrs(slot, pin, din, pout, dout, dsk, st, cn);
        BHG_CUDA(cudaEventSynchronize(c->slot_out_done[slot]));
        pool.copy(exit_pos + 3 * b, pout, (size_t)m * 24);
        pool.copy(exit_dir + 3 * b, dout, (size_t)m * 24);
        memcpy(status + b, st, (size_t)m * 4);
        if (counters) {
            memcpy(counters + b, cn, (size_t)m * 4);
            memcpy(counters + n + b, cn + m, (size_t)m * 4);
        }
        if (disk) memcpy(extras->disk_xy + 2 * b, dsk, (size_t)m * 16);
        return 0;
    };
    int si = 0;
    long long idx = 0;
    for (long long b = 0; b < n; b += chunk, si = (si + 1) % 3, idx++) {
        const long long m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t s = c->streams[si];
        // counters of a chunk live at [b, b+m) and [n + b, ...): give the kernel a chunk-local view by
        // writing attempts/accepted into a 2m block and scattering on the way back
        int32_t* cnt_chunk = counters ? d_cnt + 2 * b : nullptr;
        bhg_extras ex_chunk;
        if (disk) { ex_chunk = *extras; ex_chunk.disk_xy = d_disk + 2 * b; }
        double *pin = nullptr, *din = nullptr, *pout = nullptr, *dout = nullptr, *dsk = nullptr;
        int32_t *st = nullptr, *cn = nullptr;
        if (bounce) {
            // slot si was last used by chunk idx-3, whose results were drained at iteration idx-1
            if (idx >= 3) BHG_CUDA(cudaEventSynchronize(c->slot_in_done[si]));
            slot_ptrs(si, pin, din, pout, dout, dsk, st, cn);
            pool.copy(pin, entry_pos + 3 * b, (size_t)m * 24);
            pool.copy(din, entry_dir + 3 * b, (size_t)m * 24);
        }
        BHG_CUDA(cudaMemcpyAsync(d_pin + 3 * b, bounce ? pin : entry_pos + 3 * b, (size_t)m * 24, cudaMemcpyHostToDevice, s));
        BHG_CUDA(cudaMemcpyAsync(d_din + 3 * b, bounce ? din : entry_dir + 3 * b, (size_t)m * 24, cudaMemcpyHostToDevice, s));
        if (bounce) BHG_CUDA(cudaEventRecord(c->slot_in_done[si], s));
        rc = launch_trace(*c, d_pin + 3 * b, d_din + 3 * b, d_pout + 3 * b, d_dout + 3 * b, d_status + b, cnt_chunk,
                          nullptr, m, bhg::IN_AOS, params->image_width, params, s, disk ? &ex_chunk : nullptr);
        if (rc) return rc;
        if (disk) BHG_CUDA(cudaMemcpyAsync(bounce ? dsk : extras->disk_xy + 2 * b, d_disk + 2 * b, (size_t)m * 16, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(bounce ? pout : exit_pos + 3 * b, d_pout + 3 * b, (size_t)m * 24, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(bounce ? dout : exit_dir + 3 * b, d_dout + 3 * b, (size_t)m * 24, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(bounce ? st : status + b, d_status + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        if (counters) {
            BHG_CUDA(cudaMemcpyAsync(bounce ? cn : counters + b, cnt_chunk, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
            BHG_CUDA(cudaMemcpyAsync(bounce ? cn + m : counters + n + b, cnt_chunk + m, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        }
        if (bounce) {
            BHG_CUDA(cudaEventRecord(c->slot_out_done[si], s));
            // drain the chunk issued two iterations ago while the GPU works on the last two
            if (idx >= 2 && (rc = drain(b - 2 * chunk, (si + 1) % 3))) return rc;
        }
    }
    if (bounce) {
        const long long nchunks = idx;
        for (long long j = (nchunks >= 2 ? nchunks - 2 : 0); j < nchunks; j++)
            if ((rc = drain(j * chunk, (int)(j % 3)))) return rc;
    }
    for (auto& s : c->streams) BHG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int bhg_trace_schwarzschild_f32io(const float* entry_pos, const float* entry_dir, float* exit_pos, float* exit_dir,
                                  int32_t* status, int32_t* counters, int64_t n, const bhg_params* params,
                                  int32_t device, void* stream) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    if (n > 0 && (!entry_pos || !entry_dir || !exit_dir || !status)) return fail(BHG_ERR_INVALID_ARGUMENT, "NULL ray buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    return launch_trace(*c, (const double*)entry_pos, (const double*)entry_dir, (double*)exit_pos, (double*)exit_dir,
                        status, counters, nullptr, n, bhg::IN_AOS_F32, params->image_width, params, (cudaStream_t)stream);
}

int bhg_trace_schwarzschild_f32io_host(const float* entry_pos, const float* entry_dir, float* exit_pos, float* exit_dir,
                                       int32_t* status, int64_t n, const bhg_params* params, int32_t device) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    if (n > 0 && (!entry_pos || !entry_dir || !exit_pos || !exit_dir || !status))
        return fail(BHG_ERR_INVALID_ARGUMENT, "NULL ray buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(c->host_mu);
    const size_t vec = (size_t)n * 3 * sizeof(float);
    if ((rc = ensure_stage(c, 4 * vec + (size_t)n * sizeof(int32_t) + 1024))) return rc;
    DrainStreamsOnExit drain_streams_on_exit{c};
    char* base = (char*)c->stage;
    float* d_pin = (float*)base;
    float* d_din = (float*)(base + vec);
    float* d_pout = (float*)(base + 2 * vec);
    float* d_dout = (float*)(base + 3 * vec);
    int32_t* d_status = (int32_t*)(base + 4 * vec);
    const long long chunk = pick_chunk(n, params->image_width > 0 ? 8LL * params->image_width : 0, 1 << 19);
    int si = 0;
    for (long long b = 0; b < n; b += chunk, si = (si + 1) % 3) {
        const long long m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t s = c->streams[si];
        BHG_CUDA(cudaMemcpyAsync(d_pin + 3 * b, entry_pos + 3 * b, (size_t)m * 12, cudaMemcpyHostToDevice, s));
        BHG_CUDA(cudaMemcpyAsync(d_din + 3 * b, entry_dir + 3 * b, (size_t)m * 12, cudaMemcpyHostToDevice, s));
        rc = launch_trace(*c, (const double*)(d_pin + 3 * b), (const double*)(d_din + 3 * b), (double*)(d_pout + 3 * b),
                          (double*)(d_dout + 3 * b), d_status + b, nullptr, nullptr, m, bhg::IN_AOS_F32,
                          params->image_width, params, s);
        if (rc) return rc;
        BHG_CUDA(cudaMemcpyAsync(exit_pos + 3 * b, d_pout + 3 * b, (size_t)m * 12, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(exit_dir + 3 * b, d_dout + 3 * b, (size_t)m * 12, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(status + b, d_status + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    }
    for (auto& s : c->streams) BHG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int bhg_generate_rays_f64(const bhg_camera* cam, double r_sphere, int64_t n, double* pos, double* dir, int32_t* hit,
                          int32_t device, void* stream) {
    DeviceRestore restore_device_on_exit;
    bhg::Camera dc;
    int rc = convert_camera(cam, r_sphere, &dc);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!pos || !dir))) return fail(BHG_ERR_INVALID_ARGUMENT, "bad n or NULL ray buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    return launch_generate(*c, dc, n, pos, dir, hit, (cudaStream_t)stream);
}

int bhg_trace_camera_f64(const bhg_camera* cam, double* exit_pos, double* exit_dir, int32_t* status,
                         int32_t* counters, int64_t n, const bhg_params* params, int32_t device, void* stream) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    bhg::Camera dc;
    if ((rc = convert_camera(cam, params->r_sphere, &dc))) return rc;
    if (n > 0 && (!exit_dir || !status)) return fail(BHG_ERR_INVALID_ARGUMENT, "NULL output buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    if (n == 0) return 0;
    // the pre-pass generates each primary ray from the camera description and prepares it in one go: no ray buffer
    return launch_trace(*c, nullptr, nullptr, exit_pos, exit_dir, status, counters, nullptr, n, bhg::IN_AOS,
                        camera_image_width(dc), params, (cudaStream_t)stream, nullptr, &dc);
}

int bhg_trace_camera_f64_host(const bhg_camera* cam, double* exit_pos, double* exit_dir, int32_t* status,
                              int32_t* counters, int64_t n, const bhg_params* params, int32_t device) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    bhg::Camera dc;
    if ((rc = convert_camera(cam, params->r_sphere, &dc))) return rc;
    if (n > 0 && (!exit_dir || !status)) return fail(BHG_ERR_INVALID_ARGUMENT, "NULL output buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(c->host_mu);
    const size_t vec = (size_t)n * 3 * sizeof(double);
    if ((rc = ensure_stage(c, 2 * vec + (size_t)n * 3 * sizeof(int32_t) + 1024))) return rc;
    DrainStreamsOnExit drain_streams_on_exit{c};
    char* base = (char*)c->stage;
    double* d_pout = (double*)(base);
    double* d_dout = (double*)(base + vec);
    int32_t* d_status = (int32_t*)(base + 2 * vec);
    int32_t* d_cnt = d_status + n;
    // chunks are whole 8-row bands so that every chunk keeps the tile scheduling
    const long long chunk = pick_chunk(n, 8LL * cam->width);
    int si = 0;
    for (long long b = 0; b < n; b += chunk, si = (si + 1) % 3) {
        const long long m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t s = c->streams[si];
        bhg::Camera cc = dc;
        cc.first_ray = dc.first_ray + b;
        int32_t* cnt_chunk = counters ? d_cnt + 2 * b : nullptr;
        rc = launch_trace(*c, nullptr, nullptr, exit_pos ? d_pout + 3 * b : nullptr, d_dout + 3 * b,
                          d_status + b, cnt_chunk, nullptr, m, bhg::IN_AOS, camera_image_width(cc), params, s, nullptr, &cc);
        if (rc) return rc;
        if (exit_pos) BHG_CUDA(cudaMemcpyAsync(exit_pos + 3 * b, d_pout + 3 * b, (size_t)m * 24, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(exit_dir + 3 * b, d_dout + 3 * b, (size_t)m * 24, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(status + b, d_status + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        if (counters) {
            BHG_CUDA(cudaMemcpyAsync(counters + b, cnt_chunk, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
            BHG_CUDA(cudaMemcpyAsync(counters + n + b, cnt_chunk + m, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
        }
    }
    for (auto& s : c->streams) BHG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int bhg_trace_camera_f32_host(const bhg_camera* cam, float* exit_pos, float* exit_dir, int32_t* status, int64_t n,
                              const bhg_params* params, int32_t device) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    if (params->mode != BHG_MODE_PARITY)
        return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_trace_camera_f32_host: parity mode only");
    bhg::Camera dc;
    if ((rc = convert_camera(cam, params->r_sphere, &dc))) return rc;
    if (n > 0 && (!exit_dir || !status)) return fail(BHG_ERR_INVALID_ARGUMENT, "NULL output buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(c->host_mu);
    const size_t vec = (((size_t)n * 3 * sizeof(float)) + 255) & ~(size_t)255;
    if ((rc = ensure_stage(c, 2 * vec + (size_t)n * sizeof(int32_t) + 1024))) return rc;
    DrainStreamsOnExit drain_streams_on_exit{c};
    char* base = (char*)c->stage;
    float* d_pout = (float*)(base);
    float* d_dout = (float*)(base + vec);
    int32_t* d_status = (int32_t*)(base + 2 * vec);
    const long long chunk = pick_chunk(n, 8LL * cam->width);
    int si = 0;
    for (long long b = 0; b < n; b += chunk, si = (si + 1) % 3) {
        const long long m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t s = c->streams[si];
        bhg::Camera cc = dc;
        cc.first_ray = dc.first_ray + b;
        // FP64 integration; the trace kernel rounds the exit state to float32 as it stores it (IN_AOS_F32 store path)
        rc = launch_trace(*c, nullptr, nullptr, exit_pos ? (double*)(d_pout + 3 * b) : nullptr, (double*)(d_dout + 3 * b),
                          d_status + b, nullptr, nullptr, m, bhg::IN_AOS_F32, camera_image_width(cc), params, s, nullptr, &cc);
        if (rc) return rc;
        if (exit_pos) BHG_CUDA(cudaMemcpyAsync(exit_pos + 3 * b, d_pout + 3 * b, (size_t)m * 12, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(exit_dir + 3 * b, d_dout + 3 * b, (size_t)m * 12, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(status + b, d_status + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    }
    for (auto& s : c->streams) BHG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int bhg_sky_uv_f32(const double* exit_dir, const int32_t* status, int64_t n, float* uv, int32_t device, void* stream) {
    DeviceRestore restore_device_on_exit;
    if (n < 0 || (n > 0 && (!exit_dir || !uv))) return fail(BHG_ERR_INVALID_ARGUMENT, "bad n or NULL buffer");
    DeviceCtx* c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > c->sm_count * 16LL) blocks = c->sm_count * 16LL;
    bhg::sky_uv_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(exit_dir, status, n, (float2*)uv);
    g_launches.fetch_add(1);
    BHG_CUDA(cudaGetLastError());
    return 0;
}

int bhg_trace_camera_sky_host(const bhg_camera* cam, float* uv, int32_t* status, int64_t n, const bhg_params* params,
                              int32_t device) {
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, n);
    if (rc) return rc;
    bhg::Camera dc;
    if ((rc = convert_camera(cam, params->r_sphere, &dc))) return rc;
    if (n > 0 && (!uv || !status)) return fail(BHG_ERR_INVALID_ARGUMENT, "NULL output buffer");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(c->host_mu);
    const size_t vec = (size_t)n * 3 * sizeof(double);
    if ((rc = ensure_stage(c, vec + (size_t)n * (8 + 4) + 1024))) return rc;
    DrainStreamsOnExit drain_streams_on_exit{c};
    char* base = (char*)c->stage;
    double* d_dout = (double*)(base);
    float* d_uv = (float*)(base + vec);
    int32_t* d_status = (int32_t*)(base + vec + (size_t)n * 8);
    const long long chunk = pick_chunk(n, 8LL * cam->width);
    int si = 0;
    for (long long b = 0; b < n; b += chunk, si = (si + 1) % 3) {
        const long long m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t s = c->streams[si];
        bhg::Camera cc = dc;
        cc.first_ray = dc.first_ray + b;
        rc = launch_trace(*c, nullptr, nullptr, nullptr, d_dout + 3 * b, d_status + b, nullptr, nullptr, m,
                          bhg::IN_AOS, camera_image_width(cc), params, s, nullptr, &cc);
        if (rc) return rc;
        long long blocks = (m + 255) / 256;
        if (blocks > c->sm_count * 16LL) blocks = c->sm_count * 16LL;
        bhg::sky_uv_kernel<<<(int)blocks, 256, 0, s>>>(d_dout + 3 * b, d_status + b, m, (float2*)(d_uv + 2 * b));
        g_launches.fetch_add(1);
        BHG_CUDA(cudaGetLastError());
        BHG_CUDA(cudaMemcpyAsync(uv + 2 * b, d_uv + 2 * b, (size_t)m * 8, cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaMemcpyAsync(status + b, d_status + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    }
    for (auto& s : c->streams) BHG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

void* bhg_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        fail(BHG_ERR_OUT_OF_MEMORY, "cudaHostAlloc(%lld) failed", (long long)bytes);
        return nullptr;
    }
    return p;
}

void bhg_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int bhg_device_alloc(int64_t bytes, int32_t device, void** ptr) {
    if (!ptr || bytes <= 0) return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_device_alloc: ptr NULL or bytes <= 0");
    DeviceRestore restore_device_on_exit;
    DeviceCtx* c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    // plain cudaMalloc (not the stream-ordered pool): legacy CUDA IPC can only export such allocations
    BHG_CUDA(cudaMalloc(ptr, (size_t)bytes));
    return 0;
}

int bhg_device_free(void* ptr, int32_t device) {
    if (!ptr) return 0;
    DeviceRestore restore_device_on_exit;
    BHG_CUDA(cudaSetDevice(device));
    BHG_CUDA(cudaFree(ptr));
    return 0;
}

int bhg_ipc_export(const void* ptr, int32_t device, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "bhgeo.h documents a 64-byte handle");
    if (!ptr || !handle) return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_ipc_export: NULL argument");
    DeviceRestore restore_device_on_exit;
    BHG_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    BHG_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle, &h, sizeof(h));
    return 0;
}

int bhg_ipc_open(const unsigned char handle[64], int32_t device, void** ptr) {
    if (!ptr || !handle) return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_ipc_open: NULL argument");
    DeviceRestore restore_device_on_exit;
    DeviceCtx* c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    BHG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int bhg_ipc_close(void* ptr, int32_t device) {
    if (!ptr) return 0;
    DeviceRestore restore_device_on_exit;
    BHG_CUDA(cudaSetDevice(device));
    BHG_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

int bhg_trace_frame_shard_f64(const double* entry_pos, const double* entry_dir, int32_t entry_is_frame, int64_t m,
                              double* frame_pos, double* frame_dir, int32_t* frame_status, int64_t band_rays,
                              int64_t first_band, int64_t band_stride, const bhg_params* params, int32_t device,
                              void* stream) {
    const double* shard_pos = entry_pos;
    const double* shard_dir = entry_dir;
    DeviceRestore restore_device_on_exit;
    int rc = validate(params, m);
    if (rc) return rc;
    if (m == 0) return 0;
    if (!shard_pos || !shard_dir || !frame_pos || !frame_dir || !frame_status)
        return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_trace_frame_shard_f64: NULL buffer");
    if (band_rays < 4 || band_rays % 4 || first_band < 0 || band_stride < 1)
        return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_trace_frame_shard_f64: band_rays must be a positive multiple of 4, "
                                              "first_band >= 0, band_stride >= 1");
    DeviceCtx* c;
    if ((rc = ensure_device(device, &c))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->courier_stream) BHG_CUDA(cudaStreamCreateWithFlags(&c->courier_stream, cudaStreamNonBlocking));
    }
    const long long nb = (m + band_rays - 1) / band_rays;
    const size_t vec = (((size_t)m * 24) + 255) & ~(size_t)255, sts = (((size_t)m * 4) + 255) & ~(size_t)255;
    const size_t cnt = ((2 * (size_t)nb + 2) * 4 + 255) & ~(size_t)255;   // band_done, claimed, n_claimed, error
    char* scratch = nullptr;
    BHG_CUDA(cudaMallocAsync((void**)&scratch, 2 * vec + sts + cnt, s));
    double* o_pos = (double*)scratch;
    double* o_dir = (double*)(scratch + vec);
    int32_t* o_st = (int32_t*)(scratch + 2 * vec);
    int* band_done = (int*)(scratch + 2 * vec + sts);
    BHG_CUDA(cudaMemsetAsync(band_done, 0, cnt, s));
    // Kernels that are waited for must be resident in the context BEFORE anything spins on them: with lazy module
    // loading the first launch of a kernel may need a context-wide synchronisation, which a spinning courier would
    // never allow.  cudaFuncGetAttributes loads a kernel; once per device.
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->courier_loaded) {
            cudaFuncAttributes fa;
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::courier_kernel));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::prepare_kernel<4, bhg::IN_AOS>));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::prepare_kernel<3, bhg::IN_AOS>));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::trace_kernel<4, bhg::IN_AOS, false, false, true, false>));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::trace_kernel<3, bhg::IN_AOS, false, false, true, false>));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::cost_sample_kernel<bhg::IN_AOS>));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::cost_key_kernel<bhg::IN_AOS>));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::cost_decide_kernel));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::cost_offsets_kernel));
            BHG_CUDA(cudaFuncGetAttributes(&fa, bhg::cost_scatter_kernel));
            c->courier_loaded = true;
        }
    }
    cudaEvent_t ready, delivered;
    BHG_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    BHG_CUDA(cudaEventCreateWithFlags(&delivered, cudaEventDisableTiming));
    BHG_CUDA(cudaEventRecord(ready, s));
    bhg::CourierArgs ca;
    ca.src_pos = o_pos; ca.src_dir = o_dir; ca.src_status = o_st;
    ca.dst_pos = frame_pos; ca.dst_dir = frame_dir; ca.dst_status = frame_status;
    ca.band_done = band_done; ca.m = m; ca.band_rays = band_rays; ca.first_band = first_band; ca.band_stride = band_stride;
    ca.claimed = band_done + nb;
    ca.n_claimed = band_done + 2 * nb;
    ca.error = band_done + 2 * nb + 1;
    // SMs left to the courier: one SM sustains ~25 GB/s of peer stores (measured, profiles/r2h_courier_n*.txt), a shard
    // of 1/8 frame needs ~70 GB/s to stay hidden behind its integration
    int courier_sms = 6;
    if (const char* e = getenv("BHG_COURIER_SMS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= c->sm_count / 2) courier_sms = v;
    }
    // the trace (memset, binning, pre-pass, persistent kernel) is enqueued FIRST, then the courier on its own stream:
    // nothing the trace still needs from the driver can be held up by the courier's polling
    ShardHook hook{band_done, band_rays, courier_sms, first_band, entry_is_frame ? band_stride : 0};
    rc = launch_trace(*c, shard_pos, shard_dir, o_pos, o_dir, o_st, nullptr, nullptr, m, bhg::IN_AOS, params->image_width,
                      params, s, nullptr, nullptr, &hook);
    if (!rc) {
        BHG_CUDA(cudaStreamWaitEvent(c->courier_stream, ready, 0));
        ca.wait_limit = 200000;   // ~0.2 s without a single band completing: stop polling
        bhg::courier_kernel<<<courier_sms, 1024, 0, c->courier_stream>>>(ca);
        g_launches.fetch_add(1);
        BHG_CUDA(cudaGetLastError());
        BHG_CUDA(cudaEventRecord(delivered, c->courier_stream));
        BHG_CUDA(cudaStreamWaitEvent(s, delivered, 0));
        // sweep pass on the caller's stream, after the trace kernel and the courier: delivers every band the courier
        // did not take (none when the two ran side by side; all of them when the kernels were executed one at a time)
        ca.wait_limit = 0;
        bhg::courier_kernel<<<c->sm_count, 1024, 0, s>>>(ca);
        g_launches.fetch_add(1);
        BHG_CUDA(cudaGetLastError());
    }
    cudaFreeAsync(scratch, s);
    cudaEventDestroy(ready);
    cudaEventDestroy(delivered);
    return rc;
}

namespace {
// driver entry points without a link-time dependency on libcuda (the library must load on a machine without a driver)
typedef int (*StreamMemOp32)(void* /*CUstream*/, unsigned long long /*CUdeviceptr*/, unsigned int, unsigned int);
int stream_memop(const char* name, void* addr, int32_t value, unsigned flags, int32_t device, void* stream) {
    if (!addr || ((uintptr_t)addr & 3)) return fail(BHG_ERR_INVALID_ARGUMENT, "%s: NULL or misaligned address", name);
    DeviceRestore restore_device_on_exit;
    DeviceCtx* c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    BHG_CUDA(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(BHG_ERR_CUDA, "%s is not available in this driver", name);
    const int r = ((StreamMemOp32)fn)(stream, (unsigned long long)(uintptr_t)addr, (unsigned int)value, flags);
    if (r != 0) return fail(BHG_ERR_CUDA, "%s failed with CUresult %d", name, r);
    return 0;
}
}  // namespace

int bhg_stream_write32(void* addr, int32_t value, int32_t device, void* stream) {
    return stream_memop("cuStreamWriteValue32", addr, value, 0u /* CU_STREAM_WRITE_VALUE_DEFAULT */, device, stream);
}

int bhg_stream_wait_geq32(void* addr, int32_t value, int32_t device, void* stream) {
    return stream_memop("cuStreamWaitValue32", addr, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */, device, stream);
}

int bhg_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes, int64_t rows,
                  int32_t device, void* stream) {
    if (rows <= 0 || row_bytes <= 0) return 0;
    if (!dst || !src || dst_pitch < row_bytes || src_pitch < row_bytes)
        return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_copy_rows: NULL pointer or pitch < row_bytes");
    DeviceRestore restore_device_on_exit;
    BHG_CUDA(cudaSetDevice(device));
    BHG_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)row_bytes, (size_t)rows,
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int bhg_device_pci_bus_id(int32_t device, char* buf, int32_t len) {
    if (!buf || len < 16) return fail(BHG_ERR_INVALID_ARGUMENT, "bhg_device_pci_bus_id: buf NULL or len < 16");
    BHG_CUDA(cudaDeviceGetPCIBusId(buf, len, device));
    return 0;
}

int bhg_sum_counters(const int32_t* counters_dev, const int32_t* status_dev, int64_t n, int32_t device, void* stream,
                     int64_t* n_attempt, int64_t* n_accept, int64_t* n_integrated) {
    DeviceRestore restore_device_on_exit;
    DeviceCtx* c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    long long h[3] = {0, 0, 0};
    if (n > 0 && counters_dev && status_dev) {
        cudaStream_t s = (cudaStream_t)stream;
        BHG_CUDA(cudaMemsetAsync(c->totals, 0, 3 * sizeof(long long), s));
        int blocks = (int)((n + 1023) / 1024);
        if (blocks > c->sm_count * 8) blocks = c->sm_count * 8;
        bhg::sum_counters_kernel<<<blocks, 256, 0, s>>>(counters_dev, status_dev, n, c->totals);
        g_launches.fetch_add(1);
        BHG_CUDA(cudaGetLastError());
        BHG_CUDA(cudaMemcpyAsync(h, c->totals, sizeof(h), cudaMemcpyDeviceToHost, s));
        BHG_CUDA(cudaStreamSynchronize(s));
    }
    if (n_attempt) *n_attempt = h[0];
    if (n_accept) *n_accept = h[1];
    if (n_integrated) *n_integrated = h[2];
    return 0;
}

int64_t bhg_launch_count(void) { return g_launches.load(); }

int bhg_selftest(int32_t device, double* out8) {
    DeviceRestore restore_device_on_exit;
    DeviceCtx* c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    double* d = nullptr;
    BHG_CUDA(cudaMalloc(&d, 8 * sizeof(double)));
    BHG_CUDA(cudaMemset(d, 0, 8 * sizeof(double)));
    bhg::selftest_kernel<<<64, 256>>>(d);
    g_launches.fetch_add(1);
    BHG_CUDA(cudaGetLastError());
    double h[8];
    BHG_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    if (out8) memcpy(out8, h, sizeof(h));
    // reciprocal and tenth root to a few ulp; RHS forms agree to rounding; sincos exact by construction
    if (!(h[0] < 1e-15) || !(h[1] < 4e-15) || !(h[2] < 1e-12) || !(h[3] < 1e-15) || !(h[4] < 4e-16) || !(h[5] < 4e-15)) return fail(BHG_ERR_CUDA, "selftest out of bounds: rcp %.3e root %.3e rhs %.3e", h[0], h[1], h[2]);
    return 0;
}

int bhg_fp64_peak_tflops(int32_t device, double* tflops, double* sm_clock_mhz_est) {
    DeviceRestore restore_device_on_exit;
    DeviceCtx* c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    double* d = nullptr;
    BHG_CUDA(cudaMalloc(&d, sizeof(double) * 1024));
    cudaEvent_t e0, e1;
    BHG_CUDA(cudaEventCreate(&e0));
    BHG_CUDA(cudaEventCreate(&e1));
    const int iters = 4096;
    const int blocks = c->sm_count * 4, threads = 256;
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; rep++) {
        BHG_CUDA(cudaEventRecord(e0));
        bhg::dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001);
        g_launches.fetch_add(1);
        BHG_CUDA(cudaEventRecord(e1));
        BHG_CUDA(cudaEventSynchronize(e1));
        float ms;
        BHG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    BHG_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;  // 16 DFMA per iteration per thread
    const double tf = flops / (best_ms * 1e-3) / 1e12;
    if (tflops) *tflops = tf;
    // 64 FP64 FMA lanes per SM per clock on B200
    if (sm_clock_mhz_est) *sm_clock_mhz_est = tf * 1e12 / (2.0 * 64.0 * c->sm_count) / 1e6;
    return 0;
}

const char* bhg_last_error_string(void) { return g_err; }
int bhg_version(void) { return BHG_VERSION; }

}  // extern "C"
