/* mdkm.h -- C ABI of libmdkm.so: multi-day height-map -> XYZ unprojection + Lloyd k-means
 * on one NVIDIA B200 (sm_100a) per handle.
 *
 * This is the drop-in boundary for the ONE hot path of
 * rafael-alani/3d-point-cloud-multiday-imagery that this project accelerates
 * (SURVEY.md section 8).  The reference is pure Python and has no FFI of its own; each
 * entry point below replaces the Python/NumPy/scikit-learn code it cites, and
 * INTEGRATION.md shows the ctypes stub a maintainer would add to
 * members/rafael/disparity/plugin.py.
 *
 * Conventions
 *   - plain C, no exceptions, no torch types; every call returns MDKM_OK (0) or a negative
 *     mdkm_status; mdkm_last_error(h) gives the text for the last failure on that handle.
 *   - one handle == one CUDA device == one rank.  A handle is used by one host thread at a
 *     time (the reference calls the path from a single napari worker thread,
 *     members/rafael/disparity/widget.py:116-147).
 *   - pointers are host pointers unless the call takes a `mem` argument, in which case
 *     MDKM_MEM_DEVICE means "device pointer on the handle's device" (e.g. a torch tensor's
 *     data_ptr()).  The library never keeps a caller pointer after the call returns.
 *   - there is no CPU fallback: without a usable CUDA device mdkm_create fails.
 */
#ifndef MDKM_H_
#define MDKM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mdkm_handle mdkm_handle;

typedef enum mdkm_status {
  MDKM_OK = 0,
  MDKM_ERR_INVALID = -1,   /* bad argument */
  MDKM_ERR_CUDA = -2,      /* CUDA runtime error (see mdkm_last_error) */
  MDKM_ERR_NO_DEVICE = -3, /* no CUDA device / wrong architecture */
  MDKM_ERR_STATE = -4,     /* call order (e.g. fit before points are set) */
  MDKM_ERR_NCCL = -5,      /* NCCL missing or failed */
  MDKM_ERR_OOM = -6
} mdkm_status;

enum { MDKM_MEM_HOST = 0, MDKM_MEM_DEVICE = 1 };
/* height raster element type: float32 heights; int16 OpenCV fixed-point disparity; or the
 * reference's own "5-out-F.tif" pixel layout (disparity.py:213-224): three pixel-interleaved
 * float32 bands per pixel, band 0 = height (-disp/16), band 2 = final_defined (0 = invalid) */
enum { MDKM_HM_F32 = 0, MDKM_HM_I16 = 1, MDKM_HM_F32_GTIFF3 = 2 };
enum { MDKM_POINTS_AOS = 0, MDKM_POINTS_SOA = 1 }; /* xyzxyz... or x[n] y[n] z[n] */

#define MDKM_NCCL_UNIQUE_ID_BYTES 128
#define MDKM_IPC_HANDLE_BYTES 64

/* Library / build identification ("mdkm <ver> sm_100a"). */
const char* mdkm_version(void);

/* Create a handle on CUDA device `device`.  `cuda_stream` is an optional cudaStream_t the
 * library should enqueue on (NULL: it creates its own non-blocking stream). */
int mdkm_create(mdkm_handle** out, int device, void* cuda_stream);
void mdkm_destroy(mdkm_handle* h);
const char* mdkm_last_error(const mdkm_handle* h);

/* ---- multi-GPU: one rank per handle, NCCL over NVLink ------------------------------- */
/* Rank 0 makes the id; the caller ships the 128 bytes to the other ranks (the Python
 * host uses torch.distributed for that) and every rank calls mdkm_comm_init. */
int mdkm_comm_unique_id(unsigned char out[MDKM_NCCL_UNIQUE_ID_BYTES]);
int mdkm_comm_init(mdkm_handle* h, int n_ranks, int rank,
                   const unsigned char id[MDKM_NCCL_UNIQUE_ID_BYTES]);

/* Optional, after mdkm_comm_init on every rank (one process per GPU, one node): let the Lloyd
 * step kernel exchange the K x 4 partial sums itself, by stores into peer memory over
 * NVLink, instead of an ncclAllReduce between two kernels.  Each rank exports a CUDA IPC
 * handle of its exchange buffer (mdkm_comm_p2p_handle), the caller gathers the 64-byte
 * handles of all ranks in rank order (the Python host uses torch.distributed) and hands the
 * n_ranks * 64 bytes to mdkm_comm_p2p_open.  If the open fails (no peer access) the call
 * returns MDKM_ERR_NCCL and the NCCL path stays in use.  Collective: all ranks or none. */
int mdkm_comm_p2p_handle(mdkm_handle* h, unsigned char out[MDKM_IPC_HANDLE_BYTES]);
int mdkm_comm_p2p_open(mdkm_handle* h, const unsigned char* handles);
/* Drops the peer mappings (and this rank's exchange buffer) again and keeps the NCCL
 * communicator: the Lloyd iteration goes back to step kernel -> ncclAllReduce -> update kernel.
 * The caller uses it to make the fallback collective: when mdkm_comm_p2p_open failed on ANY
 * rank, EVERY rank calls this (no-op where nothing is open). */
int mdkm_comm_p2p_close(mdkm_handle* h);

/* The same exchange for ranks that live in ONE process (one handle per device, each driven by
 * its own host thread -- the in-process caller of the reference, widget.py:116-147): after
 * mdkm_comm_p2p_handle on every handle, mdkm_comm_p2p_buffer returns the device pointer of this
 * rank's exchange buffer, and mdkm_comm_p2p_open_ptrs takes the n_ranks pointers in rank order
 * (peer access between the devices is enabled by the call; no CUDA IPC involved). */
int mdkm_comm_p2p_buffer(mdkm_handle* h, void** out_device_ptr);
int mdkm_comm_p2p_open_ptrs(mdkm_handle* h, void* const* buffers);

/* The CUDA stream (cudaStream_t) the handle enqueues on, so that a caller that produces inputs
 * or consumes device-resident outputs on another stream can order the two with events. */
void* mdkm_get_stream(const mdkm_handle* h);

/* Tuning / measurement switches.  MDKM_OPT_SETTLE_GROUPS (default 1): 0 disables the
 * classification pass's settling of whole 128-point groups from their cached summaries, so
 * that EVERY point is read and assigned by the per-point pass in every iteration (the
 * streaming path bench.py reports its HBM fraction for); results are identical either way. */
/* MDKM_OPT_RASTER_MIRROR (default 1): 0 makes clouds that came from mdkm_unproject use the generic
 * (histogram + scatter) build of the tile-ordered mirror instead of the run-table build; a test
 * hook, the results are identical.
 * MDKM_OPT_CELL_PX (0 = automatic, 8 or 16) / MDKM_OPT_CELL_ROWS (0 = automatic, 1..64): width and
 * height, in pixels, of the x-y cells the mirror of a raster cloud is ordered by (tuning).
 * MDKM_OPT_TWO_LEVEL (-1 = automatic by cloud size, 0, 1): whether the classification pass of the
 * Lloyd kernel tests super-groups of 1024 points before it looks at 128-point groups.
 * MDKM_OPT_DEPENDENT_LAUNCH (default 1): after the first (cooperative) Lloyd kernel of a fit, the
 * following ones are launched with programmatic stream serialisation -- the next iteration's kernel
 * is set up while the current one drains -- instead of the cooperative attribute (the runtime does
 * not combine the two); 0 = every launch cooperative.  Results are identical. */
enum { MDKM_OPT_SETTLE_GROUPS = 1, MDKM_OPT_RASTER_MIRROR = 2, MDKM_OPT_CELL_PX = 3, MDKM_OPT_CELL_ROWS = 4,
       MDKM_OPT_TWO_LEVEL = 5, MDKM_OPT_DEPENDENT_LAUNCH = 6 };
int mdkm_set_option(mdkm_handle* h, int option, long long value);

/* ---- K1: unprojection ----------------------------------------------------------------
 * Replaces members/rafael/disparity/plugin.py:148 (h = -disp/16), :151-152 (validity:
 * isfinite & |h| <= max_abs & mask), :157-160 ((y,x)=np.where(valid), P=[x,y,z]) and the
 * (absent in the reference) multi-day concatenation, for the pixel range
 * [pix_begin, pix_begin+pix_count) of the flattened [D,H,W] stack.  `hm` / `mask` point at
 * pixel `pix_begin`.  hm_dtype MDKM_HM_F32: heights, used as is (hm_scale ignored);
 * MDKM_HM_I16: OpenCV fixed-point disparity, h = hm_scale * disp (the reference uses
 * -1/16); MDKM_HM_F32_GTIFF3: 12 bytes per pixel, band 2 acts as the validity mask (ANDed
 * with `mask` when that is given too).  mask may be NULL.  detrend != 0 additionally applies the per-day plane fit of
 * plugin.py:161-171 (z becomes the signed distance to the day's least-squares plane); it
 * needs whole days in the range.  The points stay resident in the handle, in np.where
 * order (day-major, row-major); *n_points_out receives how many there are on this rank. */
int mdkm_unproject(mdkm_handle* h, const void* hm, int hm_dtype, float hm_scale,
                   const uint8_t* mask, int D, int H, int W, int64_t pix_begin,
                   int64_t pix_count, float max_abs, int detrend, int mem,
                   int64_t* n_points_out);

/* Optional: have the NEXT mdkm_unproject also stream the cloud it produces into host memory
 * (float32 [n,3]; (z,y,x) columns when napari_order != 0, i.e. plugin.py:192's
 * `points_coords`), slab by slab while later slabs are still being uploaded and unprojected.
 * out_host must hold capacity_points >= pix_count points (the valid count is not known in
 * advance) and should be page-locked.  The binding is consumed by that one call; the array
 * is complete after mdkm_wait (or any later call that touches the cloud). */
int mdkm_bind_cloud_output(mdkm_handle* h, float* out_host, int64_t capacity_points,
                           int napari_order);

/* Load an already unprojected cloud (this rank's shard).  Replaces the `X` argument of the
 * reference's KMeans(...).fit_predict(X) call (members/jasraj/land_use_classification/
 * core.py:227-228) for d = 3. */
int mdkm_set_points(mdkm_handle* h, const float* xyz, int64_t n, int layout, int mem);

/* Number of resident points on this rank / over all ranks of the communicator (the second
 * one is a collective the first time it is called for a cloud). */
int64_t mdkm_num_points(const mdkm_handle* h);
int64_t mdkm_num_points_global(mdkm_handle* h);

/* Fetch m points by index as float32 [m,3] (x,y,z) into host memory (what `X[seeds]` does for
 * init="random", sklearn/cluster/_kmeans.py:1014-1021).  With a communicator the indices
 * are global (shards concatenated in rank order), the call is collective and every rank
 * receives all m points. */
int mdkm_gather_points(mdkm_handle* h, const int64_t* idx, int m, float* out_xyz);

/* Copy the resident cloud out as float32 [n,3].  napari_order != 0 gives (z,y,x) columns as
 * plugin.py:192 builds `points_coords`; 0 gives (x,y,z). */
int mdkm_get_cloud(mdkm_handle* h, float* out, int napari_order, int mem);

/* Same, but returns as soon as the copy is enqueued: it proceeds on a side stream while
 * later calls on this handle (mdkm_fit) compute.  `out` (page-locked host memory for a real
 * overlap) holds the cloud once mdkm_wait -- or any later call that touches the cloud --
 * has returned. */
int mdkm_get_cloud_async(mdkm_handle* h, float* out, int napari_order, int mem);
int mdkm_wait(mdkm_handle* h);

/* Segments of the resident cloud: one per day of the unprojected pixel range (the reference
 * treats every stereo pair / day separately, plugin.py:106), one for mdkm_set_points.
 * mdkm_segment_offsets writes n_segments + 1 point offsets (the last one is n). */
int mdkm_num_segments(const mdkm_handle* h);
int mdkm_segment_offsets(const mdkm_handle* h, int64_t* out);

/* Ground-levelling of plugin.py:181-192, per segment (day): h_min / h_max =
 * np.percentile(z, 2) / np.percentile(z, 98) with numpy's linear interpolation (exact order
 * statistics by radix select on the device); z -= h_min in the resident cloud;
 * height_norm_out (optional, n floats, host or device per `mem`) =
 * clip((z - h_min) / (h_max - h_min + 1e-6), 0, 1), the 'height' colour property.
 * h_min_out / h_max_out: optional arrays of n_segments doubles.  Rank-local: needs whole days
 * on this rank (shard the stack by days).  Call it before mdkm_fit. */
int mdkm_ground_level(mdkm_handle* h, float* height_norm_out, int mem, double* h_min_out,
                      double* h_max_out);

/* ---- K2..K4: Lloyd k-means -----------------------------------------------------------
 * Replaces sklearn.cluster.KMeans(n_clusters=k, init=<array>, n_init=1,
 * algorithm="lloyd", max_iter, tol).fit(X) as the reference calls it (core.py:227-228;
 * algorithm in sklearn/cluster/_kmeans.py:630-758, 1436-1563 and _k_means_lloyd.pyx).
 * init: k x 3 float64, row-major, same coordinates as the points.
 * labels_out: int32[n_local] (may be NULL); centroids_out: float64[k*3]; n_iter_out,
 * inertia_out as scikit-learn's n_iter_ / inertia_.  With a communicator the point set is
 * the union of all ranks' shards and every rank gets the same centroids / n_iter /
 * inertia and its own shard's labels.  No host synchronisation per iteration. */
int mdkm_fit(mdkm_handle* h, int k, const double* init, int max_iter, double tol,
             int32_t* labels_out, int labels_mem, double* centroids_out, int* n_iter_out,
             double* inertia_out);

/* Diagnostics of the last mdkm_fit / mdkm_lloyd_step on this rank:
 * n_refined = point-iterations whose FP32 distances were within the rounding-error bound
 * of a tie and were re-decided in float64; n_relocations = empty-cluster relocations. */
int mdkm_fit_stats(const mdkm_handle* h, int64_t* n_refined, int64_t* n_relocations,
                   double* tol_scaled);

/* More diagnostics of the last mdkm_fit on this rank: *worklist_groups = 128-point groups that
 * went through the per-point pass, summed over the iterations of the fit (fused iterations
 * only); *groups = groups of the cloud.  1 - worklist_groups / (groups * n_iter) is the share
 * of group-iterations settled from the cached summaries without reading a point. */
int mdkm_fit_worklist(const mdkm_handle* h, int64_t* worklist_groups, int64_t* groups);

/* One E-step + M-step sums with the given centroids (test hook for single-step parity;
 * mirrors sklearn.cluster._k_means_lloyd.lloyd_iter_chunked_dense up to the sums,
 * _k_means_lloyd.pyx:23-152).  sums_out: float64[k*3] = sum of member coordinates,
 * counts_out: int64[k]; both GLOBAL over ranks when a communicator is set. */
int mdkm_lloyd_step(mdkm_handle* h, int k, const double* centroids, int32_t* labels_out,
                    int labels_mem, double* sums_out, int64_t* counts_out);

/* Labels of the resident points for given centroids and the resulting inertia: the E-step
 * alone, i.e. KMeans.predict / -KMeans.score (sklearn/cluster/_kmeans.py:742-756, 1068-1095;
 * inertia: _k_means_common.pyx:94-124).  labels_out may be NULL; inertia is global over ranks. */
int mdkm_predict(mdkm_handle* h, int k, const double* centroids, int32_t* labels_out,
                 int labels_mem, double* inertia_out);

/* k-means++ seeding on the device (sklearn/cluster/_kmeans.py:180-278).  The random draws
 * of numpy.random.RandomState stay on the host and are supplied by the caller in
 * scikit-learn's order: `first_index` is the result of RandomState.choice(n) for the first
 * centre, then for each further centre `n_local_trials` uniforms in [0,1) from rand_vals
 * (row-major [(k-1), n_local_trials], n_local_trials <= 16; scikit-learn uses
 * 2 + int(ln k)); the device does the distances (float64), the potentials, the
 * cumulative-sum search and the greedy choice among the trials, all k centres without a
 * host round trip.  centers_out: float64[k*3]; indices_out: int64[k] (may be NULL).
 * With a communicator the cloud is the concatenation of the ranks' shards in rank order:
 * first_index and the returned indices are GLOBAL, every rank passes the same arguments and
 * gets the same centres (the ranks' sums are exchanged with NCCL, three small all-reduces
 * per centre). */
int mdkm_kmeans_plusplus(mdkm_handle* h, int k, int64_t first_index, const double* rand_vals,
                         int n_local_trials, double* centers_out, int64_t* indices_out);

/* The first fit on a cloud builds two acceleration structures that later fits on the same
 * cloud reuse: the tile-ordered mirror of the points and the per-group summaries (box +
 * fixed-point sums).  mdkm_drop_caches discards them, so that the next fit pays for them again
 * (bench.py times every fit that way). */
int mdkm_drop_caches(mdkm_handle* h);

/* Timing aid for bench.py: when enabled, CUDA events bracket every batch of back-to-back
 * launches of the Lloyd step kernel inside mdkm_fit; mdkm_profile_read returns the summed
 * duration and the number of launches it covers and resets both (average launch duration =
 * their ratio; batches that ended early on convergence make it a lower bound). */
int mdkm_profile_enable(mdkm_handle* h, int on);
int mdkm_profile_read(mdkm_handle* h, double* step_kernel_ms, int* n_step_launches,
                      int* n_kernel_launches_total);

/* The same for the other kernel groups of the path (CUDA events on the handle's stream around
 * the kernels only -- host work and copies between them are outside): summed duration and work
 * count since the last read of that phase.  count = pixels (UNPROJECT: count + scan + scatter
 * kernels of mdkm_unproject), points (BUILD: tile-ordered mirror + group summaries of a cold
 * fit; FINAL: the final labelling / inertia pass of mdkm_fit), launches (STEP). */
enum { MDKM_PHASE_UNPROJECT = 0, MDKM_PHASE_BUILD = 1, MDKM_PHASE_STEP = 2, MDKM_PHASE_FINAL = 3,
       MDKM_PHASE_COUNT = 4 };
int mdkm_profile_phase(mdkm_handle* h, int phase, double* ms, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* MDKM_H_ */
