/* pst_b200.h -- C ABI of libpst_b200.so, the B200 (sm_100a) genotype hot path.
 *
 * The reference (PySnpTools) reaches its native code through nine symbols of the third-party
 * `bed_reader` extension.  Each entry point below names the reference call site it
 * replaces.  Conventions, identical for every function:
 *   - plain pointers and sizes only; no ownership transfer; the caller frees what it allocated;
 *   - return value 0 = success, non-zero = failure with the message in pstb_last_error()
 *     (thread-local, valid until the next call on that thread);
 *   - `d_` pointers are device (HBM) pointers of the current CUDA device, `h_` pointers are host
 *     pointers; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions taking only `d_` pointers are asynchronous on `stream`; `_host` functions return
 *     after the result is in the caller's host buffer;
 *   - dtype codes: PSTB_F32 / PSTB_F64 / PSTB_I8; order codes: PSTB_ORDER_F (iid fastest) /
 *     PSTB_ORDER_C (sid fastest).
 *
 * Packed genotype store in HBM: `d_packed` holds sid_count records of `ld` bytes each; record j is
 * the PLINK SNP-major record of SNP j (ceil(iid_count/4) bytes, 4 genotypes per byte, least
 * significant pair first; 3-byte file header stripped).  `ld >= ceil(iid_count/4)`; a 16-byte
 * multiple enables the 128-bit load path (pstb_packed_ld gives the preferred value).
 *
 * Index selection on either axis: `idx` (uint32 device vector of `n` entries) or, when idx == NULL,
 * the arithmetic progression start + k*step, k < n.  The host layer resolves negative / boolean /
 * nested indexers (pysnptools/pstreader/_subset.py:55-142) before calling.
 */
#ifndef PST_B200_H
#define PST_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { PSTB_F32 = 0, PSTB_F64 = 1, PSTB_I8 = 2 };
enum { PSTB_ORDER_F = 0, PSTB_ORDER_C = 1 };
/* standardize modes */
enum { PSTB_STD_NONE = 0, PSTB_STD_UNIT = 1, PSTB_STD_BETA = 2 };

typedef struct pstb_axis {
    const uint32_t* idx; /* device pointer or NULL */
    int64_t start;       /* used when idx == NULL */
    int64_t step;        /* used when idx == NULL */
    int64_t n;           /* number of selected entries */
} pstb_axis;

/* ---- housekeeping ------------------------------------------------------------------------- */
int pstb_version(void);
const char* pstb_last_error(void);
/* replaces bed_reader.get_num_threads (standardizer.py:110, util/__init__.py:335): the GPU path has
 * no host thread pool; returns the SM count of the current device (0 if no device). */
int pstb_sm_count(void);
int64_t pstb_packed_ld(int64_t iid_count);
/* number of this library's kernels launched by the calling process so far (bench.py "gpu_launches") */
int64_t pstb_launch_count(void);
/* page-locked host memory for the `_host` entry points: buffers from here are copied to / from the GPU
 * directly (asynchronously, overlapped with the kernels); ordinary pageable buffers work too but go
 * through an internal pinned staging ring.  NULL on failure. */
void* pstb_host_alloc(int64_t bytes);
int pstb_host_free(void* p);
/* the `_host` entry points keep their device / staging buffers (chunk ring, workspace, the device copy of K) cached per
 * calling thread so that repeated calls do not pay cudaMalloc / cudaFree; this returns them to the driver. */
int pstb_host_release(void);
/* multi-GPU hosts: bind the CALLING thread (and the threads it creates later: this library's copy workers) to the CPUs of the NUMA
 * node `device` hangs off and prefer that node's memory, so that pinned buffers allocated afterwards and the staging copies stay
 * local to the GPU's PCIe root.  One process per GPU calls it once before its first `_host` call (the reference has no such notion:
 * its readers run on whatever cores the OS picks).  Returns the node (>= 0), -1 when the host exposes no NUMA information for the
 * device (nothing changed), -2 on error (message in pstb_last_error). */
int pstb_numa_bind(int device);

/* ---- K1: decode  (replaces open_bed(...).read -> Rust read_f32/f64/i8; bed.py:337-343) ------ */
int pstb_decode(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count,
                pstb_axis iid, pstb_axis sid, int count_a1,
                void* d_out, int dtype, int order, void* stream);

/* ---- K2: fused decode + per-SNP statistics + standardize ------------------------------------
 * replaces read (bed.py:337-343) followed by standardize_f32/f64 (standardizer.py:114,120) without
 * materialising the raw matrix.  mode = PSTB_STD_UNIT | PSTB_STD_BETA (a, b = Beta parameters).
 * use_stats == 0: d_stats[n_sid][2] (float64, C order: mean, std) is WRITTEN; use_stats != 0: it is
 * READ (UnitTrained / BetaTrained; unittrained.py:47-70, betatrained.py:47-63).  d_out may be NULL to
 * compute statistics only.  dtype is PSTB_F32 or PSTB_F64. */
int pstb_decode_standardize(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count,
                            pstb_axis iid, pstb_axis sid, int count_a1,
                            int mode, double a, double b, int use_stats, double* d_stats,
                            void* d_out, int dtype, int order, void* stream);

/* ---- K2f: standardize an existing float matrix in place -------------------------------------
 * replaces standardize_f32 / standardize_f64(snps, is_beta, a, b, apply_in_place, use_stats, stats,
 * num_threads) (standardizer.py:109-121).  d_val is [n_iid, n_sid] in `order`; d_stats as above.
 * d_work: device scratch of pstb_standardize_work_bytes(n_sid) bytes. */
int64_t pstb_standardize_work_bytes(int64_t n_sid);
int pstb_standardize(void* d_val, int dtype, int order, int64_t n_iid, int64_t n_sid,
                     int mode, double a, double b, int apply_in_place, int use_stats, double* d_stats,
                     void* d_work, void* stream);

/* ---- gather: replaces subset_f64_f64 / subset_f32_f64 / subset_f32_f32 (util/__init__.py:341-375)
 * out[i, j, k] = in[rows[i], cols[j], k] for 3-D [n, m, v] arrays, C or F contiguous each. */
int pstb_subset(const void* d_in, int dtype_in, int order_in, int64_t n_in, int64_t m_in, int64_t v,
                pstb_axis rows, pstb_axis cols, void* d_out, int dtype_out, int order_out, void* stream);

/* ---- K0: pack  (replaces to_bed -> Rust write_f32/f64/i8; bed.py:300-314) --------------------
 * d_val [n_iid, n_sid] (any dtype / order) -> d_packed (ld bytes per SNP).  *d_bad (int32 on device)
 * is set non-zero when a value is not one of 0, 1, 2, NaN (-127 for int8). */
int pstb_pack(const void* d_val, int dtype, int order, int64_t n_iid, int64_t n_sid, int count_a1,
              uint8_t* d_packed, int64_t ld, int32_t* d_bad, void* stream);

/* ---- K3: kinship  K = X X^T on tcgen05 -------------------------------------------------------
 * replaces the block loop of SnpReader._read_kernel (snpreader.py:651-655: read + standardize +
 * val.dot(val.T) + `K +=`; snpdata.py:203-206).
 *
 * pstb_kernel_workspace_bytes: scratch needed for n_iid selected individuals and SNP chunks of
 *   `chunk` SNPs (fp16 hi/lo operand planes + per-SNP tables).
 * pstb_snp_kernel: K (float32, [n_iid, n_iid], ld = n_iid, both triangles filled) =
 *   sum over the selected SNPs of x_j x_j^T, where x_j is the standardized column of SNP j.
 *   accumulate != 0 adds to the existing lower triangle of d_K before mirroring (multi-call
 *   streaming).  d_stats [n_sid][2] float64 is written (or read when use_stats).  Missing genotypes
 *   contribute 0 (mean imputation, standardizer.py:145-163) -- also under the Identity standardizer (statistics (0, 1) passed with
 *   use_stats), where the reference's val.dot(val.T) would propagate NaN (snpdata.py:203-206): the one deliberate deviation, see
 *   INTEGRATION.md section 2.  Exact-dosage 2-term tensor-core path
 *   (see `low_term` below), fp32 accumulation in tensor memory: relative Frobenius error vs float64
 *   ~1e-6 (fp16 low term) / (2..5)e-6 (fp8 low term). */
int64_t pstb_kernel_workspace_bytes(int64_t n_iid, int64_t chunk);
/* Precision / speed of the exact-dosage path: K = L B^T with L = g - mu' exact in fp16 (missing -> fp16(mu - mu'), so that mean
 * imputation is exact whatever the missing pattern) and B = w m (g - mu) = hi + lo in fp16.  The low term L lo^T only has to
 * carry 4-5 bits, so it can run on the fp8 pipe (e4m3 x e4m3, half the tensor cycles): relative Frobenius error (2..5)e-6
 * instead of ~1e-6 (north_star gate: 1e-5), ~23 % faster on cfg3.  Every kernel entry point takes the mode PER CALL
 * (`low_term`), so concurrent callers cannot race on it:
 *   PSTB_LOW_TERM_FP16 / PSTB_LOW_TERM_FP8: as named;
 *   PSTB_LOW_TERM_AUTO: fp8 when the call multiplies at least as many SNPs as individuals (4 x as many for Beta, whose weights
 *     concentrate on the rare SNPs) and at least 256 -- the error estimate is statistical.  A caller that shards the SNPs of
 *     one kernel over several calls / GPUs knows the global SNP count and passes FP8 / FP16 explicitly;
 *   PSTB_LOW_TERM_DEFAULT: the process-wide default, initially AUTO (environment PSTB_SYRK_FP8LO=0/1 overrides it), which
 *     pstb_set_syrk_low_term changes (returns the previous default; an invalid mode only reports). */
enum { PSTB_LOW_TERM_DEFAULT = -1, PSTB_LOW_TERM_FP16 = 0, PSTB_LOW_TERM_FP8 = 1, PSTB_LOW_TERM_AUTO = 2 };
int pstb_set_syrk_low_term(int mode);
int pstb_snp_kernel(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count,
                    pstb_axis iid, pstb_axis sid, int count_a1,
                    int mode, double a, double b, int use_stats, double* d_stats,
                    float* d_K, int accumulate, int mirror,
                    void* d_work, int64_t work_bytes, int64_t chunk, int low_term, void* stream);
/* K-tile sharding for kernels that do not fit one GPU (BASELINE cfg5: 500 000 iids -> K = 1 TB): the lower triangle is cut
 * into 256 x 256 tiles, rank r of `world` owns every world-th tile of a fixed rasterisation and computes them for ALL
 * selected SNPs from its own copy of the packed store -- no reduction, no collective.
 * pstb_kernel_tile_count / pstb_kernel_tile_coords: how many tiles `rank` owns and their (I, J) block coordinates
 * (rows [256 I, +256), columns [256 J, +256), J <= I; host array int32 [count][2]).
 * pstb_snp_kernel_tiles: like pstb_snp_kernel, but writes d_tiles [count][256][256] float32 (each owned tile stored whole,
 * diagonal tiles with both triangles; entries beyond iid.n are left untouched). */
int64_t pstb_kernel_tile_count(int64_t n_iid, int rank, int world);
int pstb_kernel_tile_coords(int64_t n_iid, int rank, int world, int32_t* h_ij);
int pstb_snp_kernel_tiles(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count,
                          pstb_axis iid, pstb_axis sid, int count_a1,
                          int mode, double a, double b, int use_stats, double* d_stats,
                          float* d_tiles, int rank, int world, int accumulate,
                          void* d_work, int64_t work_bytes, int64_t chunk, int low_term, void* stream);
/* Expand compact tiles (the layout above, tiles of `rank` of `world`) into the full symmetric K [n_iid, n_iid] (both triangles;
 * entries of other ranks' tiles are left untouched).  The SNP-sharded multi-GPU path accumulates into compact tiles
 * (rank 0 of world 1 = the whole lower triangle), all-reduces them -- half the bytes of the square matrix -- and expands. */
int pstb_kernel_from_tiles(const float* d_tiles, int64_t n_iid, int rank, int world, float* d_K, void* stream);
/* Bands of a compact-tile kernel, for overlapping the NCCL reduction of the SNP-sharded multi-GPU path with its last SNP chunk
 * ("overlappable by reducing finished tile-rows early"): pstb_snp_kernel_tiles_band multiplies ONE chunk of SNPs
 * (sid.n <= chunk) for tiles [tile_begin, tile_end) of `rank`'s list only, leaving `reserve_sms` SMs idle for a concurrent collective.
 * flags bit 0 = first band of the chunk: statistics and operand planes are built into d_work; later bands (flags = 0) reuse them.
 * low_term must be PSTB_LOW_TERM_FP16 or _FP8.  pstb_kernel_from_tiles_range expands tiles [tile_begin, tile_end) into the square K. */
int pstb_snp_kernel_tiles_band(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count,
                               pstb_axis iid, pstb_axis sid, int count_a1,
                               int mode, double a, double b, int use_stats, double* d_stats,
                               float* d_tiles, int rank, int world, int accumulate,
                               void* d_work, int64_t work_bytes, int64_t chunk, int low_term,
                               int64_t tile_begin, int64_t tile_end, int flags, int reserve_sms, void* stream);
int pstb_kernel_from_tiles_range(const float* d_tiles, int64_t n_iid, int rank, int world, int64_t tile_begin, int64_t tile_end,
                                 float* d_K, const double* d_u, void* stream);
/* Deferred rank-one part (multi-GPU): with bit 1 of `accumulate` (pstb_snp_kernel_tiles) or of `flags` (pstb_snp_kernel_tiles_band) set,
 * the call leaves the float64 vector v [n_iid] of the exact-dosage path (K_ik = sum_j L_ij B_kj + v_k) in its workspace instead of adding
 * it to the tiles; pstb_kernel_workspace_rank1 returns its address.  The caller sums the vectors of its calls, all-reduces the total
 * with the tiles (n_iid doubles) and passes it as d_u to pstb_kernel_from_tiles_range, which adds it while expanding (NULL: nothing). */
double* pstb_kernel_workspace_rank1(void* d_work, int64_t n_iid, int64_t chunk);
/* PSTB_LOW_TERM_DEFAULT / _AUTO resolved for a kernel of `total_sid` SNPs (returns PSTB_LOW_TERM_FP16 or _FP8): a caller that splits one
 * kernel over several calls resolves the mode once and passes the explicit value to each. */
int pstb_resolve_low_term(int low_term, int64_t total_sid, int64_t n_iid, int mode);
/* Train x test kernel (what FaST-LMM builds from SnpKernel + the *Trained standardizers: unittrained.py:47-70,
 * betatrained.py:47-63 applied to a second iid set, then train.val.dot(test.val.T)):
 *   d_out [n_r, n_c] float32, C order (ld = n_c):  out[i, k] (+)= sum_j x_ij y_kj
 * x = the `_r` (row / train) selection, y = the `_c` (column / test) selection, possibly of two different stores; both sides
 * select the same number of SNPs (position j of sid_r pairs with position j of sid_c).  Both are standardized with ONE set of
 * per-SNP statistics: use_stats == 0 computes them from the ROW side and writes d_stats [n_sid][2]; use_stats != 0 reads them.
 * Missing genotypes contribute 0 (mean imputation).  Same exact-dosage tensor-core path as pstb_snp_kernel (the row side is the
 * exact left operand, the column side the weighted right operand). */
int64_t pstb_cross_kernel_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t chunk);
int pstb_snp_cross_kernel(const uint8_t* d_packed_r, int64_t ld_r, int64_t iid_count_r, int64_t sid_count_r,
                          pstb_axis iid_r, pstb_axis sid_r, int count_a1_r,
                          const uint8_t* d_packed_c, int64_t ld_c, int64_t iid_count_c, int64_t sid_count_c,
                          pstb_axis iid_c, pstb_axis sid_c, int count_a1_c,
                          int mode, double a, double b, int use_stats, double* d_stats,
                          float* d_out, int accumulate, void* d_work, int64_t work_bytes, int64_t chunk, int low_term, void* stream);
/* K = V V^T for a float matrix V [n_iid, n_sid] already in HBM (float32 / float64, C or F order): replaces the
 * val.dot(val.T) of SnpData._read_kernel (snpdata.py:203-206).  Same fp16 hi/lo tensor-core path and workspace. */
int pstb_float_kernel(const void* d_val, int dtype, int order, int64_t n_iid, int64_t n_sid, float* d_K, int accumulate,
                      int mirror, void* d_work, int64_t work_bytes, int64_t chunk, void* stream);
/* ---- K3 in float64 -----------------------------------------------------------------------------
 * The reference computes a kernel in the dtype the caller asks for: val.dot(val.T) (snpdata.py:203-206) is a DGEMM for the default
 * dtype=float64, and its unit tests compare float64 kernels to 10 decimals (kernelreader/test.py:48-50, :189; test.py:535-553).
 * These entry points are that contract on the GPU: the same fused decode + exact-count statistics + standardize, into a float64
 * panel, followed by an fp64-FMA SYRK on the CUDA cores (two-level summation: one register partial per SNP chunk, then K += it).
 * ~1e-13 relative to the reference's float64 path; ~100 x its CPU rate (B200 has full-rate fp64).  d_K is float64 [n_iid, n_iid].
 * pstb_kernel_f64_workspace_bytes: one float64 panel of `chunk` SNPs.  `chunk` need not be a multiple of 64 here. */
int64_t pstb_kernel_f64_workspace_bytes(int64_t n_iid, int64_t chunk);
int pstb_snp_kernel_f64(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count,
                        pstb_axis iid, pstb_axis sid, int count_a1,
                        int mode, double a, double b, int use_stats, double* d_stats,
                        double* d_K, int accumulate, int mirror,
                        void* d_work, int64_t work_bytes, int64_t chunk, void* stream);
/* K = V V^T in float64 for a float64 matrix V [n_iid, n_sid] (C or F order) already in HBM. */
int pstb_float_kernel_f64(const double* d_val, int order, int64_t n_iid, int64_t n_sid, double* d_K, int accumulate, int mirror,
                          void* stream);
int pstb_mirror_lower_f64(double* d_K, int64_t n, int64_t ldk, void* stream);
/* Test/bench hook for the tensor-core stage alone: K_lower (+)= (hi+lo)(hi+lo)^T minus lo*lo^T, on
 * fp16 planes [n_pad, k_pad] (row-major, k_pad % 64 == 0, n_pad % 128 == 0). */
int pstb_syrk_planes(const void* d_hi, const void* d_lo, int64_t n, int64_t n_pad, int64_t k_pad,
                     float* d_K, int64_t ldk, int accumulate, float out_scale, void* stream);
/* mirror the lower triangle into the upper one (in place) */
int pstb_mirror_lower(float* d_K, int64_t n, int64_t ldk, void* stream);
/* float32 K -> float64 / float32 copy with optional scale (KernelData dtype contract, kernelreader.py:245-302) */
int pstb_convert_kernel(const float* d_K, int64_t n, void* d_out, int dtype, double scale, void* stream);

/* ---- host-buffer entry points (what a bed_reader-style binding calls with NumPy arrays) ------ */
/* read_f32/f64/i8 equivalent: h_packed = file bytes after the 3-byte header (tight: ceil(iid_count/4)
 * bytes per SNP); h_iid_idx / h_sid_idx: int64 host vectors or NULL (= all); h_out: caller-allocated
 * [n_iid, n_sid] host array.  mode != PSTB_STD_NONE fuses the standardize (h_stats [n_sid][2] float64). */
int pstb_read_host(const uint8_t* h_packed, int64_t iid_count, int64_t sid_count,
                   const int64_t* h_iid_idx, int64_t n_iid, const int64_t* h_sid_idx, int64_t n_sid,
                   int count_a1, int mode, double a, double b, int use_stats, double* h_stats,
                   void* h_out, int dtype, int order);
/* SnpReader._read_kernel on host buffers (snpreader.py:623-668: the loop of read + standardize + val.dot(val.T) + `K +=`):
 * h_packed as for pstb_read_host; h_K: caller-allocated [n_iid, n_iid] float32 / float64 (C order; K is symmetric), both
 * triangles filled; h_stats [n_sid][2] float64 written (read when use_stats).  The packed records are streamed to the GPU in
 * slices overlapped with the tensor-core work; chunk = SNPs per operand-plane chunk (a multiple of 64).  For large kernels
 * (n_iid >= 8192, >= 8 chunks) the copy-out of K is overlapped too: the last chunks are multiplied band-major from the bottom
 * band of tiles up and finished row ranges of K leave while the bands above multiply (PSTB_HOST_KERNEL_OVERLAP=0: plain loop). */
int pstb_snp_kernel_host(const uint8_t* h_packed, int64_t iid_count, int64_t sid_count,
                         const int64_t* h_iid_idx, int64_t n_iid, const int64_t* h_sid_idx, int64_t n_sid,
                         int count_a1, int mode, double a, double b, int use_stats, double* h_stats,
                         void* h_K, int dtype, int64_t chunk, int low_term);
/* the same loop in float64 arithmetic (pstb_snp_kernel_f64): h_K is float64 [n_iid, n_iid]; chunk a multiple of 64. */
int pstb_snp_kernel_host_f64(const uint8_t* h_packed, int64_t iid_count, int64_t sid_count,
                             const int64_t* h_iid_idx, int64_t n_iid, const int64_t* h_sid_idx, int64_t n_sid,
                             int count_a1, int mode, double a, double b, int use_stats, double* h_stats,
                             double* h_K, int64_t chunk);
/* standardize_f32/f64 equivalent on a host array (H2D, K2f, D2H). */
int pstb_standardize_host(void* h_val, int dtype, int order, int64_t n_iid, int64_t n_sid,
                          int mode, double a, double b, int apply_in_place, int use_stats, double* h_stats);
/* subset_* equivalent on host arrays. */
int pstb_subset_host(const void* h_in, int dtype_in, int order_in, int64_t n_in, int64_t m_in, int64_t v,
                     const int64_t* h_rows, int64_t n_rows, const int64_t* h_cols, int64_t n_cols,
                     void* h_out, int dtype_out, int order_out);

#ifdef __cplusplus
}
#endif
#endif /* PST_B200_H */
