/* bhgeo.h — C ABI of the B200-native batched Schwarzschild null-geodesic tracer.
 *
 * Drop-in boundary for ONE path of bldevries/blackhole_geodesic_calculator: the per-ray call into
 * curvedpy's geodesic solver.  The reference has no FFI layer of its own (it is three Blender add-on
 * scripts calling a pure-Python package); each entry point below names the reference interface it
 * replaces (paths relative to /root/reference).  INTEGRATION.md shows the ctypes binding a maintainer
 * adds on the reference side.
 *
 * Conventions: geometrised units G = c = 1, horizon r_s = 2 M (raytracer/RelativisticRenderEngine.py:95);
 * positions are relative to the black-hole centre (RelativisticRenderEngine.py:278,
 * LimitedRelativisticRenderEngine.py:265); directions are coordinate tangents, normalised on output
 * (RelativisticRenderEngine.py:371-372 needs |d_z| <= 1).  All buffers are caller-owned.  Functions return
 * 0 on success and a negative bhg_error on failure (never throw); bhg_last_error_string() describes the
 * last failure on the calling thread.  Thread-safe: calls may come from any host thread.
 */
#ifndef BHGEO_H
#define BHGEO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHG_VERSION 120 /* 0.2.0: + bhg_trace_frame_shard_f64, bhg_stream_write32 / _wait_geq32, bhg_trace_camera_f32_host
                           * (0.1.1: bhg_params.reserved became bhg_params.coords, same layout) */

/* per-ray status codes written to `status` */
enum bhg_status {
    BHG_ESCAPED = 0,           /* left the sphere of influence: exit state is on r = r_sphere            */
    BHG_CAPTURED = 1,          /* crossed r_s + eps_horizon: result['hit_blackhole'] (RRE.py:297, LIM.py:308) */
    BHG_START_INSIDE_HOLE = 2, /* entry radius <= r_s + eps_horizon: result['start_inside_hole'] (RRE.py:296) */
    BHG_LAMBDA_EXHAUSTED = 3,  /* affine length lambda_max reached first: mes['error']=='Outside' (LIM.py:311-312);
                                  the normal ending of the RRE fixed-length call (RRE.py:293-294,307-308)   */
    BHG_STEP_FAILED = 4,       /* step size underflow / non-finite state (scipy status -1)                */
    BHG_MISSED_SPHERE = 5      /* camera entry points only: the primary ray never meets the sphere of influence
                                  (the flat `scene.ray_cast` finds no "isBH" hit, LIM.py:224-237); exit_dir is
                                  the unchanged flat direction, exit_pos is NaN                              */
};

enum bhg_error {
    BHG_OK = 0,
    BHG_ERR_INVALID_ARGUMENT = -1,
    BHG_ERR_CUDA = -2,
    BHG_ERR_NO_DEVICE = -3,
    BHG_ERR_OUT_OF_MEMORY = -4
};

/* integration algorithm */
enum bhg_mode {
    BHG_MODE_PARITY = 0, /* 8-state spherical system + scipy RK45 control: reproduces the reference path */
    BHG_MODE_PLANE = 1   /* optional: 6-state integration in each ray's conserved orbital plane           */
};

/* Chart of the Cartesian positions / directions at the boundary.  The integration itself always runs in the
 * spherical Schwarzschild chart of README.md:162-172 (the current curvedpy generation: GeodesicIntegratorSchwarzschild
 * + Conversions().convert_xyz_to_sph, RelativisticRenderEngine.py:134,289-291), x = r sin(th) cos(ph) etc.
 * BHG_COORDS_ISOTROPIC maps entry states from, and results back to, the isotropic Cartesian chart
 * (r = rho (1 + r_s / 4 rho)^2; radial tangent component x (1 - a^2), tangential x (1 + a)^2, a = r_s / 4 rho): the
 * chart of the older solver generation `SchwarzschildGeodesic` ("uses the Schwarzschild metric in cartesian
 * coordinates", README.md:174; LimitedRelativisticRenderEngine.py:90,273-278) - README Fig. 5 and Fig. 6, the only
 * known answers the reference holds for this path, are reproduced to the pixel in this chart and in no other
 * (tests/golden/readme_fig5_fig6.npz, tests/test_readme_figures.py).  eps_horizon stays an offset in the
 * Schwarzschild radius; lambda_max counts affine length of the unit Schwarzschild-chart tangent. */
enum bhg_coords {
    BHG_COORDS_SCHWARZSCHILD = 0,
    BHG_COORDS_ISOTROPIC = 1
};

/* memory layout of the ray buffers */
enum bhg_layout {
    BHG_LAYOUT_SOA = 0, /* in: 6 planes of n doubles px,py,pz,dx,dy,dz; out: 6 planes px,py,pz,dx,dy,dz  */
    BHG_LAYOUT_AOS = 1  /* in: pos[n][3], dir[n][3]; out: pos[n][3], dir[n][3] (numpy [N,3] arrays)       */
};

/* Solver parameters.  Replaces the keyword arguments of curvedpy's calc_trajectory / ray_trace at
 * RelativisticRenderEngine.py:293-294 and LimitedRelativisticRenderEngine.py:273-278 plus the constructor
 * argument `mass` (RelativisticRenderEngine.py:134).  rtol/atol default to scipy's 1e-3 / 1e-6 because no
 * reference engine passes them. */
typedef struct bhg_params {
    double M;           /* black-hole mass, r_s = 2 M                                      */
    double r_sphere;    /* radius of the sphere of influence; +inf = no outer event (RRE)  */
    double rtol;        /* RK45 relative tolerance                                         */
    double atol;        /* RK45 absolute tolerance                                         */
    double max_step;    /* maximum affine step (+inf = unbounded; RRE.py:57-60 maps -1 to inf) */
    double eps_horizon; /* capture event at r = r_s + eps_horizon                          */
    double lambda_max;  /* affine-length bound (curve_end, RRE.py:61,294); <= 0 selects 10 r_sphere */
    int32_t mode;       /* enum bhg_mode                                                   */
    int32_t refill_threshold; /* warp work queue: idle lanes that trigger a refill, 1..32; 0 = adaptive (default):
                                 refill once the idle lane-iterations since the last refill reach a fixed budget */
    int32_t image_width;      /* coherence hint: the rays are a row-major image (or stack of images) of this
                                 width, as the reference's s -> y -> x loop produces them (RRE.py:195-218); the
                                 queue then hands every warp a 4 x 8 pixel tile instead of 32 pixels of one row.
                                 0 = no hint.  Ignored unless image_width % 4 == 0 and n % (8 image_width) == 0.
                                 Scheduling only: results are bit-identical with and without the hint.       */
    int32_t coords;           /* enum bhg_coords: chart of every position / direction / radius that crosses this
                                 boundary (entry and exit states, r_sphere, disk radii and hit points, polyline
                                 samples).  0 = Schwarzschild coordinates (default).                          */
} bhg_params;

/* Fills *p with the reference defaults (M=1, r_sphere=60, rtol=1e-3, atol=1e-6, max_step=inf,
 * eps_horizon=0.01, lambda_max=0 -> auto, parity mode). */
void bhg_default_params(bhg_params* p);

/* Batched trace on DEVICE buffers (what torch / the benchmark call with tensor.data_ptr()).
 * Replaces N calls of GeoInt.calc_trajectory (RelativisticRenderEngine.py:293-294) or SW.ray_trace
 * (LimitedRelativisticRenderEngine.py:273-278) and the end-state extraction at
 * RelativisticRenderEngine.py:296-310 / LimitedRelativisticRenderEngine.py:308-319.
 *   layout SOA: in = 6*n doubles (planes), out = 6*n doubles (planes)
 *   layout AOS: in = pos[n][3] then dir[n][3] given as two pointers via in/in_dir; same for out.
 *   in_dir/out_dir are ignored (may be NULL) for SOA.
 *   status   : n int32 (enum bhg_status)
 *   counters : optional (NULL ok) 2*n int32: [0,n) RK45 step attempts, [n,2n) accepted steps
 *   order    : optional (NULL ok) n int32 permutation: queue slot k processes ray order[k]
 *   device   : CUDA device ordinal; stream: cudaStream_t (NULL = default stream).  Asynchronous on `stream`. */
int bhg_trace_schwarzschild_f64(const double* in, const double* in_dir, double* out, double* out_dir,
                                int32_t* status, int32_t* counters, const int32_t* order, int64_t n,
                                int32_t layout, const bhg_params* params, int32_t device, void* stream);

/* Optional in-flight products of the same integration.  disk_xy (n x 2 doubles, NaN = no
 * hit) receives the first crossing of the equatorial plane z = 0 whose radius lies in [disk_r_in, disk_r_out]:
 * the continuous form of checkHitDisk's polyline scan (LimitedRelativisticRenderEngine.py:283-302,413-438),
 * located on the dense output like the terminal events; crossings after capture / exit do not count.
 * Within ONE accepted RK step only the first plane crossing is examined (a step spans far less than half an orbit in
 * theta at the default tolerances; checkHitDisk's scan of a sampled polyline has the same resolution limit at its
 * sample spacing).  disk_r_in / disk_r_out must satisfy 0 <= disk_r_in <= disk_r_out (BHG_ERR_INVALID_ARGUMENT).
 * Parity mode only.  Never changes exit_pos / exit_dir / status. */
typedef struct bhg_extras {
    double disk_r_in, disk_r_out; /* same length unit as M; the disk is off unless disk_r_out > 0 and disk_xy != NULL */
    double* disk_xy;
    /* Polyline: positions sampled at lambda_j = linspace(0, lambda_max, poly_n)[j] up to the termination time - what
     * curvedpy returns for nr_points_curve (RelativisticRenderEngine.py:293-294,299-300) and what checkHitDisk scans
     * (LimitedRelativisticRenderEngine.py:283-285).  poly_xyz: n x poly_n x 3 doubles (entries at and beyond
     * poly_count[i] are left untouched), poly_count: n int32.  Off unless poly_n >= 2 and both pointers are set.
     * Needs an explicit params->lambda_max > 0.  Parity mode, float64 AOS layout. */
    int32_t poly_n;
    int32_t reserved;
    double* poly_xyz;
    int32_t* poly_count;
} bhg_extras;

/* bhg_trace_schwarzschild_f64 / _host with extras (extras == NULL is identical to the plain call). */
int bhg_trace_schwarzschild_f64_ex(const double* in, const double* in_dir, double* out, double* out_dir,
                                   int32_t* status, int32_t* counters, const int32_t* order, int64_t n,
                                   int32_t layout, const bhg_params* params, const bhg_extras* extras,
                                   int32_t device, void* stream);
int bhg_trace_schwarzschild_f64_host_ex(const double* entry_pos, const double* entry_dir, double* exit_pos,
                                        double* exit_dir, int32_t* status, int32_t* counters, int64_t n,
                                        const bhg_params* params, const bhg_extras* extras, int32_t device);

/* Same, on HOST [N,3] arrays (numpy, Blender's bundled Python): stages H2D, traces, stages D2H, and returns
 * when the results are in the host buffers.  Pinned buffers (bhg_host_alloc) are copied asynchronously in
 * overlapping chunks.  counters may be NULL. */
int bhg_trace_schwarzschild_f64_host(const double* entry_pos, const double* entry_dir, double* exit_pos,
                                     double* exit_dir, int32_t* status, int32_t* counters, int64_t n,
                                     const bhg_params* params, int32_t device);

/* float32 I/O variants: entry_pos / entry_dir / exit_pos / exit_dir are [n][3] float32 (Blender's mathutils vectors
 * and image buffers are float32: RelativisticRenderEngine.py:181-182,223).  Inputs are widened exactly to FP64, the
 * integration is the same FP64 algorithm, results are rounded once to float32.  28 B/ray cross PCIe instead of 100.
 * _f32io: device buffers, asynchronous on `stream` (exit_pos may be NULL); _f32io_host: host buffers (pinned for full
 * speed), returns when the results are in place. */
int bhg_trace_schwarzschild_f32io(const float* entry_pos, const float* entry_dir, float* exit_pos, float* exit_dir,
                                  int32_t* status, int32_t* counters, int64_t n, const bhg_params* params,
                                  int32_t device, void* stream);
int bhg_trace_schwarzschild_f32io_host(const float* entry_pos, const float* entry_dir, float* exit_pos, float* exit_dir,
                                       int32_t* status, int64_t n, const bhg_params* params, int32_t device);

/* Pinhole camera of the reference's render loop: replaces the per-pixel direction construction of
 * RelativisticRenderEngine.ray_trace (RRE.py:185-189,195-230: loop order s -> y -> x, pixel offsets, jitter,
 * rotation by the camera matrix, normalisation) and the flat-space hit on the "isBH" sphere
 * (LimitedRelativisticRenderEngine.py:224,265).  Ray i of a call is ray first_ray + i of that loop order.
 * jitter = 1 draws the sub-pixel offsets from Philox-4x32-10(counter = ray index, key = seed), a counter-based
 * stand-in for the reference's sequential random.random() stream (RRE.py:189,227). */
typedef struct bhg_camera {
    double origin[3];    /* camera position relative to the black-hole centre (RRE.py:278)       */
    double rotation[9];  /* row-major camera-to-world rotation; camera looks along local -z       */
    double fov_x, fov_y; /* field_of_view_x / _y (RRE.py:72-73,224-225)                           */
    int64_t first_ray;
    uint64_t seed;       /* sampling_seed (RRE.py:189)                                            */
    int32_t width, height;
    int32_t jitter;      /* 0: pixel centres, 1: Philox                                           */
    int32_t reserved;    /* must be 0 */
} bhg_camera;

/* Generates n primary rays on the device: pos[n][3] = entry point on the sphere |p| = r_sphere (NaN if the ray
 * misses), dir[n][3] = unit direction, hit[n] (NULL ok) = 0 or BHG_MISSED_SPHERE.  Device buffers.
 * The trace entry points treat a NaN entry position as "missed": status BHG_MISSED_SPHERE, exit_dir = the
 * unchanged input direction, exit_pos = NaN. */
int bhg_generate_rays_f64(const bhg_camera* cam, double r_sphere, int64_t n, double* pos, double* dir, int32_t* hit,
                          int32_t device, void* stream);

/* Generate + trace: the whole curved-spacetime part of one frame (or tile) from a 176-byte camera description;
 * no ray buffer crosses PCIe.  (A streaming generator kernel writes the rays to stream-ordered device scratch,
 * then the trace kernel runs: measured faster than generating inside the latency-bound trace kernel.)
 * Device output buffers, AoS [n][3]; exit_pos may be NULL when
 * the consumer needs directions only (CamEdition.py:228 reads ray_end[...,3:6]; RRE.py:246 uses end_dir only).
 * Rays that miss the sphere get BHG_MISSED_SPHERE.  params->image_width is set from the camera automatically. */
int bhg_trace_camera_f64(const bhg_camera* cam, double* exit_pos, double* exit_dir, int32_t* status,
                         int32_t* counters, int64_t n, const bhg_params* params, int32_t device, void* stream);

/* Same with HOST output buffers (pinned or pageable): traces in chunks and overlaps the D2H of finished chunks
 * with the integration of the next ones.  exit_pos and counters may be NULL. */
int bhg_trace_camera_f64_host(const bhg_camera* cam, double* exit_pos, double* exit_dir, int32_t* status,
                              int32_t* counters, int64_t n, const bhg_params* params, int32_t device);

/* Same with float32 host outputs: FP64 integration, the exit state rounded once to float32 as it is stored (Blender's
 * mathutils vectors are float32, RelativisticRenderEngine.py:181-182).  exit_pos may be NULL: exit_dir + status are
 * what the RRE / CAM consumers read (RelativisticRenderEngine.py:246, RelativisticRenderEngineCamEdition.py:228) -
 * 16 bytes per ray come back over PCIe instead of 52.  Parity mode only. */
int bhg_trace_camera_f32_host(const bhg_camera* cam, float* exit_pos, float* exit_dir, int32_t* status, int64_t n,
                              const bhg_params* params, int32_t device);

/* Sky-lookup coordinates of exit directions: replaces the arithmetic of background_hit (RRE.py:366-378,
 * LIM.py:383-408): uv[i] = (-atan2(d_y,d_x)/pi, 2 (1 - acos(d_z)/pi) - 1) as float32 pairs, NaN for rays whose
 * status is captured / start-inside / failed (the reference paints those black without a lookup); status may
 * be NULL (all rays mapped).  exit_dir is AoS [n][3].  Device buffers, asynchronous on `stream`. */
int bhg_sky_uv_f32(const double* exit_dir, const int32_t* status, int64_t n, float* uv, int32_t device, void* stream);

/* Fused camera -> sky coordinates with HOST outputs: generate, trace, map, and copy back only uv[n][2] float32 and
 * status[n] (12 B/ray instead of 52): everything the host needs to composite a background-only frame. */
int bhg_trace_camera_sky_host(const bhg_camera* cam, float* uv, int32_t* status, int64_t n, const bhg_params* params,
                              int32_t device);

/* Pinned host memory for the staging path (optional convenience). */
void* bhg_host_alloc(int64_t bytes);
void bhg_host_free(void* p);

/* Peer-mapped device memory for the multi-GPU frame gather (no reference counterpart: the reference is one
 * process; its only parallelism is the offline mp.Pool camera pre-run, RelativisticRenderEngineCamEdition.py:216).
 * The rank that owns a frame allocates its exit buffers with bhg_device_alloc, exports a 64-byte CUDA IPC handle,
 * the other ranks (one process per GPU) open it and pass the mapped pointers as out / out_dir / status of
 * bhg_trace_schwarzschild_f64 together with `order` = their ray indices: every GPU then stores its exit states
 * straight into the owner's HBM over NVLink while it integrates, and no gather follows.
 * bhg_ipc_open must be called from a different process than the exporter (CUDA restriction). */
int bhg_device_alloc(int64_t bytes, int32_t device, void** ptr);
int bhg_device_free(void* ptr, int32_t device);
int bhg_ipc_export(const void* ptr, int32_t device, unsigned char handle[64]);
int bhg_ipc_open(const unsigned char handle[64], int32_t device, void** ptr);
int bhg_ipc_close(void* ptr, int32_t device);
/* Asynchronous strided copy on `stream` (cudaMemcpy2DAsync, device to device, copy engine): `rows` rows of
 * `row_bytes`, source pitch `src_pitch`, destination pitch `dst_pitch`; dst may be a bhg_ipc_open mapping.  Used to
 * deal a rank's bands of exit states into the frame owner's buffers while the next piece is still integrating. */
int bhg_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes, int64_t rows,
                  int32_t device, void* stream);

/* One rank's shard of a frame that lives in another GPU's memory.  The shard is every band_stride-th band of
 * `band_rays` consecutive rays, starting at band first_band: `m` rays in all (the frame's last band may be shorter).
 * entry_is_frame = 1: entry_pos / entry_dir are the WHOLE frame's float64 AoS arrays (every rank holds them, as every
 * Blender process would) and the shard's bands are read in place; 0: they hold the m rays of the shard, compacted.
 * The exit states are delivered to the same band positions of frame_pos / frame_dir / frame_status (bhg_ipc_open
 * mappings of the owner's buffers, or local pointers on the owner itself) WHILE the integration runs: the trace
 * kernel leaves a few SMs to a courier kernel that carries every completed band as contiguous 16-byte vectors over
 * NVLink.  One trace launch per shard, no pieces, no gather afterwards; asynchronous on `stream`.  band_rays must be a
 * multiple of 4 (with params->image_width = W: bands of 8 image rows, band_rays = 8 W, keep the tile scheduling).
 * Replaces nothing in the reference (single process); it is the SURVEY 8(e) gather to the rank that owns the frame. */
int bhg_trace_frame_shard_f64(const double* entry_pos, const double* entry_dir, int32_t entry_is_frame, int64_t m,
                              double* frame_pos, double* frame_dir, int32_t* frame_status, int64_t band_rays,
                              int64_t first_band, int64_t band_stride, const bhg_params* params, int32_t device,
                              void* stream);

/* Stream-ordered 32-bit flag in device memory (driver stream memory operations, executed by the GPU front end: they
 * need no SM, so they progress while a persistent trace kernel owns every register file).  `addr` may be a
 * bhg_ipc_open mapping: the ranks of a sharded frame post "my shard has arrived" into the frame owner's memory with
 * bhg_stream_write32 after their trace kernel and the owner's stream waits for all of them with
 * bhg_stream_wait_geq32 - no collective, no host synchronisation.  wait: (int32)*addr - value >= 0. */
int bhg_stream_write32(void* addr, int32_t value, int32_t device, void* stream);
int bhg_stream_wait_geq32(void* addr, int32_t value, int32_t device, void* stream);

/* PCI bus id of `device` ("0000:1b:00.0"), so that a host process can place its pinned frame buffers on the NUMA
 * node the GPU hangs off (api.pinned_empty(..., device=)); buf needs >= 16 bytes. */
int bhg_device_pci_bus_id(int32_t device, char* buf, int32_t len);

/* Totals of the last completed trace on `device` from the calling thread's point of view: sum of RK45
 * attempts and of RHS evaluations (nfev = 2 + 6 attempts per integrated ray) — used for roofline
 * accounting.  Only valid if the trace was given a `counters` buffer; otherwise returns zeros. */
int bhg_sum_counters(const int32_t* counters_dev, const int32_t* status_dev, int64_t n, int32_t device,
                     void* stream, int64_t* n_attempt, int64_t* n_accept, int64_t* n_integrated);

/* Number of kernels this library has launched in this process (all threads). */
int64_t bhg_launch_count(void);

/* Device-side numerical self-test of the FP64 building blocks (reciprocal, inverse tenth root, RHS against
 * IEEE division form, table sincos against the library).  Writes 8 doubles of max errors to out8 (NULL ok);
 * returns 0 if all are within bounds. */
int bhg_selftest(int32_t device, double* out8);

/* FP64 FMA throughput microbenchmark (dependent-free DFMA chains, all SMs), in TFLOP/s; used as the
 * measured roofline denominator because MEASURED_PEAKS.json carries no FP64 entry. */
int bhg_fp64_peak_tflops(int32_t device, double* tflops, double* sm_clock_mhz_est);

const char* bhg_last_error_string(void);
int bhg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BHGEO_H */
