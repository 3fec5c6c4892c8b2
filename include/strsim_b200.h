/*
 * strsim_b200.h -- C ABI of the B200-native row-wise string similarity library
 * (libpolars_strsim_b200.so).  Plain pointers and sizes only; no torch / C++ types.
 *
 * What each group replaces in the reference (paths relative to /root/reference):
 *
 *   _polars_plugin_<measure>, _polars_plugin_field_<measure>, _polars_plugin_get_last_error_message,
 *   _polars_plugin_get_version
 *       the symbols that `#[polars_expr(output_type=Float64)]` generates for the five functions in
 *       src/expressions/mod.rs:8-31 (pyo3-polars-derive 0.11 / polars-ffi 0.43.1, Cargo.lock:588-589,
 *       855-856,874-875).  Polars dlopen()s the shared object found in the package directory
 *       (polars_strsim/__init__.py:11-16) and calls these; no CPython API is involved.
 *
 *   strsim_b200_compute_arrow / strsim_b200_compute_host
 *       parallel_apply(), src/expressions/strsim.rs:41-107: validation (dtype, length / scalar
 *       broadcast), chunk alignment, null propagation (polars-core arity kernels, call sites
 *       strsim.rs:63,66,68,88,92,96) and evaluation of one measure over every row.  Host buffers in,
 *       host buffers out; H2D / kernels / D2H happen inside.
 *
 *   strsim_b200_column_upload / strsim_b200_compute_device
 *       the same evaluation with both columns already resident in HBM (what bench.py times as
 *       `value`); results stay on the device.
 *
 * Measures (SimilarityFunctionType, src/expressions/strsim.rs:9-15) and their arithmetic
 * (strsim.rs:125-345) are implemented by hand-written sm_100a kernels; there is no CPU fallback:
 * every entry point fails with STRSIM_ERR_CUDA when no CUDA device is usable.
 *
 * Thread safety: all entry points are re-entrant; calls on different host threads run on separate
 * CUDA streams (the plugin is registered is_elementwise=True, polars_strsim/__init__.py:15, so the
 * engine may call it concurrently from several threads).
 */
#ifndef STRSIM_B200_H
#define STRSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define STRSIM_API __declspec(dllexport)
#else
#define STRSIM_API __attribute__((visibility("default")))
#endif

/* SimilarityFunctionType, src/expressions/strsim.rs:9-15 */
enum strsim_measure {
    STRSIM_LEVENSHTEIN = 0,   /* strsim.rs:125-162 */
    STRSIM_JARO = 1,          /* strsim.rs:180-245 */
    STRSIM_JARO_WINKLER = 2,  /* strsim.rs:257-272 */
    STRSIM_JACCARD = 3,       /* strsim.rs:286-308 */
    STRSIM_SORENSEN_DICE = 4  /* strsim.rs:322-345 */
};

enum strsim_status {
    STRSIM_OK = 0,
    STRSIM_ERR_SHAPE = 1,    /* "Inputs must have the same length, or one of them must be a Utf8
                                literal." (strsim.rs:48-52) */
    STRSIM_ERR_DTYPE = 2,    /* input is not a String (Utf8View / Utf8 / LargeUtf8) column
                                (`.str()?`, strsim.rs:46-47) */
    STRSIM_ERR_ARGUMENT = 3, /* NULL pointer, unknown measure, ... */
    STRSIM_ERR_CUDA = 4,     /* no device / launch or copy failure: there is NO CPU fallback */
    STRSIM_ERR_NOMEM = 5
};

/* One chunk of an Arrow Utf8View / BinaryView column, exactly as laid out by the Arrow C Data
 * Interface (SURVEY.md 8(b) "Input layout"): buffers[0] validity, buffers[1] views,
 * buffers[2..2+V) data buffers.  `offset` applies to views AND validity. */
typedef struct strsim_view_chunk {
    const void *views;               /* 16-byte views, at least offset+length of them */
    const uint8_t *validity;         /* LSB-first bitmap or NULL (= all valid) */
    int64_t offset;                  /* ArrowArray.offset */
    int64_t length;                  /* ArrowArray.length */
    const void *const *data_buffers; /* V variadic data buffers (may be NULL when V == 0) */
    const int64_t *data_buffer_sizes;/* V sizes in bytes */
    int64_t n_data_buffers;          /* V */
} strsim_view_chunk;

/* Number of int32 debug integers per row written when `dbg_ints` is non-NULL:
 * [flag, la, lb, x0, x1, x2] with the convention of oracle/strsim_oracle.c (flag 0 general,
 * 1 byte-equal, 2 one side empty, 3 Jaro single-char; lev x0=d; jaro x0=m,x1=t; jw +x2=l;
 * jaccard x0=inter,x1=union; dice x0=inter,x1=la+lb).  Used by the parity tests. */
#define STRSIM_DBG_INTS 6

/* ---- host-buffer entry point (end-to-end path) ---------------------------------------------------
 * a / b: the chunks of the two columns (their chunkings may differ).  Total lengths must be equal,
 * or either may be 1 (scalar broadcast, strsim.rs:61-66,85-92; both orientations return
 * max(len) rows -- the reference's 1-row result for a literal on the LEFT, strsim.rs:73, is a bug
 * that is deliberately not reproduced).
 * out_values: n_rows doubles.  out_validity: ceil(n_rows/8) bytes, LSB-first, or NULL if the
 * caller does not want it; bit = both inputs valid (README.md:69-70).  Values under null rows are
 * 0.0.  out_null_count may be NULL.  Host pointers may be pageable or pinned (pinned memory is
 * DMA'd directly; pageable memory goes through an internal pinned staging ring). */
STRSIM_API int strsim_b200_compute_host(int measure, const strsim_view_chunk *a, size_t n_a_chunks,
                                        const strsim_view_chunk *b, size_t n_b_chunks,
                                        double *out_values, uint8_t *out_validity,
                                        int64_t *out_null_count, int32_t *dbg_ints);

/* Same, for several measures over ONE upload of the two columns (the five README expressions,
 * README.md:47-51, evaluate the same pair of columns five times; the PCIe transfer dominates the
 * end-to-end cost).  out_values[m] / dbg_ints[m] receive measure measures[m];
 * out_validity / out_null_count are shared (the null mask does not depend on the measure).
 * n_measures <= 8.  Downloads of finished measures overlap the kernels of the next one. */
STRSIM_API int strsim_b200_compute_host_multi(const int *measures, size_t n_measures,
                                              const strsim_view_chunk *a, size_t n_a_chunks,
                                              const strsim_view_chunk *b, size_t n_b_chunks,
                                              double *const *out_values, uint8_t *out_validity,
                                              int64_t *out_null_count, int32_t *const *dbg_ints);

/* Host call that can keep what it uploads (the five README expressions are five
 * separate plugin calls over the same two columns, and end to end the PCIe upload dominates).  Like
 * strsim_b200_compute_host_multi, but each column is given EITHER as host chunks OR as `resident_x`,
 * a column already in HBM (then x / n_x are ignored), and with keep_x != NULL the column this call
 * uploaded -- pipelined with the kernels, as always -- is handed to the caller (*keep_x, to be freed
 * with strsim_b200_column_free) instead of being dropped.  The Polars plugin entry points use this
 * with a small cache keyed by the Arrow buffer addresses of the input arrays they own. */
typedef struct strsim_b200_column strsim_b200_column;
STRSIM_API int strsim_b200_compute_host_keep(const int *measures, size_t n_measures,
                                             const strsim_view_chunk *a, size_t n_a_chunks,
                                             const strsim_b200_column *resident_a,
                                             strsim_b200_column **keep_a,
                                             const strsim_view_chunk *b, size_t n_b_chunks,
                                             const strsim_b200_column *resident_b,
                                             strsim_b200_column **keep_b,
                                             double *const *out_values, uint8_t *out_validity,
                                             int64_t *out_null_count, int32_t *const *dbg_ints);

/* ---- Arrow C Data Interface entry point -----------------------------------------------------------
 * Inputs are BORROWED (not released).  Accepted formats: "vu" (Utf8View); "u"/"U" (Utf8/LargeUtf8) are
 * converted to views on the host first; dictionary-encoded arrays of those (integer index format with
 * `ArrowSchema.dictionary` set).  Binary layouts are a dtype error, as in the reference (strsim.rs:46-47).  `out` receives a Float64 array
 * (format "g") with a release callback; the caller owns it. */
struct ArrowArray;
struct ArrowSchema;
STRSIM_API int strsim_b200_compute_arrow(int measure, const struct ArrowSchema *a_schema,
                                         const struct ArrowArray *const *a_chunks, size_t n_a_chunks,
                                         const struct ArrowSchema *b_schema,
                                         const struct ArrowArray *const *b_chunks, size_t n_b_chunks,
                                         struct ArrowArray *out);

/* ---- device-resident columns ------------------------------------------------------------------------
 * Upload copies views, validity and data buffers to HBM once (zero host-side repacking: the Arrow
 * buffers are copied as they are).  Compute is asynchronous on `stream` (a cudaStream_t passed as
 * void*; NULL = the library's per-thread stream) except that rows which do not fit the short-string
 * kernels are finished by follow-up kernels after one internal stream synchronisation.
 * d_out_values: device pointer, n_rows doubles.  d_out_validity: device pointer, ceil(n_rows/32)*4
 * bytes or NULL.  d_dbg_ints: device pointer, n_rows*STRSIM_DBG_INTS int32 or NULL. */
STRSIM_API int strsim_b200_column_upload(const strsim_view_chunk *chunks, size_t n_chunks,
                                         strsim_b200_column **out);
/* Dictionary-encoded input (Arrow dictionary / Polars Categorical; SURVEY.md 8(f).4 -- the reference rejects
 * such columns at `.str()?`, strsim.rs:46-47, so this widens the drop-in rather than mirroring it).  One
 * chunk = the indices of the rows + the dictionary they index (one chunk of views, any of the String
 * layouts after the plugin's conversion).  Only the indices and the dictionary cross PCIe; the device
 * materialises each row's view from the dictionary (dict_gather_kernel) and the measures run on the result
 * as on any other resident column.  A row is null when its index is null, out of range, or names a null
 * dictionary entry. */
typedef struct strsim_dict_chunk {
    strsim_view_chunk values;   /* the dictionary */
    const void *indices;        /* ArrowArray.buffers[1] of the encoded chunk */
    int32_t index_bytes;        /* 1, 2, 4 or 8 */
    int32_t index_signed;       /* Arrow int8..int64 (1) or uint8..uint32 (0) */
    const uint8_t *validity;    /* of the indices, LSB-first, or NULL */
    int64_t offset, length;     /* ArrowArray.offset / .length of the encoded chunk */
} strsim_dict_chunk;
STRSIM_API int strsim_b200_column_upload_dictionary(const strsim_dict_chunk *chunks, size_t n_chunks,
                                                    strsim_b200_column **out);
STRSIM_API void strsim_b200_column_free(strsim_b200_column *col);
STRSIM_API int64_t strsim_b200_column_length(const strsim_b200_column *col);
/* bytes of HBM the column occupies, and the device it lives on */
STRSIM_API int64_t strsim_b200_column_device_bytes(const strsim_b200_column *col);
STRSIM_API int strsim_b200_column_device(const strsim_b200_column *col);
/* algorithmic bytes of the column as SURVEY.md 8(d) counts them: 16 B/row of views + out-of-line
 * payload of rows with byte length > 12 (+ validity bits); filled in at upload */
STRSIM_API int64_t strsim_b200_column_algorithmic_bytes(const strsim_b200_column *col);
/* Re-runs the column-statistics pre-pass over a resident column on `stream` (NULL = the library's
 * per-thread stream) and waits for it: OR / AND of every string byte (stats_views_kernel over the views'
 * inline bytes, stats_bytes_kernel over the data buffers), which choose the kernel instantiation (5 / 6 /
 * 7 bit planes for ASCII columns, the general launches otherwise).  An upload runs the same kernels once;
 * bench.py calls this to time the pre-pass next to the measure kernel ("all kernels of the
 * measure incl. pre-pass").  No counterpart in the reference (its per-row decode looks at every byte
 * anyway, strsim.rs:131-140). */
STRSIM_API int strsim_b200_column_restat(strsim_b200_column *col, void *stream);
STRSIM_API int strsim_b200_compute_device(int measure, const strsim_b200_column *a,
                                          const strsim_b200_column *b, double *d_out_values,
                                          uint32_t *d_out_validity, int32_t *d_dbg_ints, void *stream);

/* Several DISTINCT measures (n_measures <= 5) over the same resident columns in ONE fused pass
 * (the five README expressions, README.md:47-51, evaluate the same two columns):
 * views and bytes are read once, the position mask of every character is computed once and feeds
 * the Myers column, the Jaro match step and the multiset step together; Jaro / Jaro-Winkler share
 * the match and transposition counts, Jaccard / Sorensen-Dice the intersection.  d_out_values[k] and
 * d_dbg_ints[k] (array or entries may be NULL) belong to measures[k]; results are bit-identical to
 * the single-measure calls. */
STRSIM_API int strsim_b200_compute_device_multi(const int *measures, size_t n_measures,
                                                const strsim_b200_column *a,
                                                const strsim_b200_column *b,
                                                double *const *d_out_values,
                                                uint32_t *d_out_validity,
                                                int32_t *const *d_dbg_ints, void *stream);

/* ---- housekeeping -------------------------------------------------------------------------------- */
/* device used by the calling thread's subsequent calls (default: STRSIM_B200_DEVICE env or 0) */
STRSIM_API int strsim_b200_set_device(int device);
STRSIM_API int strsim_b200_device_count(void);
/* device the calling thread's calls use right now (-1: no usable device) */
STRSIM_API int strsim_b200_get_device(void);
/* One host / plugin call over several GPUs (north_star 4; the reference fans one call out over the workers of
 * Polars' pool and re-assembles the chunks in the same call, strsim.rs:72-104).  With the environment
 * variable STRSIM_B200_DEVICES = "0,1,2,3" | "0-7" | "all", every host call of at least 65536 rows is cut
 * into one contiguous row range per listed device -- split_offsets (strsim.rs:21-39) on units of 64 rows,
 * so that no validity byte is shared -- and each range is uploaded (its views and the stretch of the data
 * buffers its rows reference), computed and downloaded by a worker thread of its device into its range of
 * the caller's buffers.  No collective, no peer traffic.  Columns kept by such a call
 * (strsim_b200_compute_host_keep, the plugin's cache) are sharded the same way and
 * strsim_b200_column_device() reports 1000 + number of shards for them.
 * strsim_b200_shard_cuts writes the n_shards + 1 boundaries of that partition (no GPU needed). */
STRSIM_API void strsim_b200_shard_cuts(int64_t n_rows, int n_shards, int64_t *cuts);
/* Multi-GPU hosts: binds the calling thread (and the threads and pinned buffers it creates from now on,
 * first-touch) to the CPUs of the NUMA node the device's PCIe link hangs off, so that host<->device
 * copies do not cross the socket interconnect.  Returns the node, or -1 when there is nothing to do
 * (one node, no topology information, or the node's CPUs are outside the thread's affinity mask).
 * The reference has no counterpart: rayon's pool is node-agnostic (strsim.rs:72-73). */
STRSIM_API int strsim_b200_bind_thread_near_device(int device);
/* Polars plugin calls keep recently uploaded input columns in HBM (and hold the Arrow arrays they own
 * alive, so that an address can only ever mean the same bytes): at most STRSIM_B200_CACHE_BYTES of
 * HBM (default 4 GiB), 8 columns, dropped after STRSIM_B200_CACHE_TTL seconds without use (default 10; a
 * reaper thread enforces it in an idle process too); STRSIM_B200_CACHE=0 disables it.
 * cache_clear drops everything now; cache_stats: [0] hits, [1] misses, [2] columns held, [3] bytes */
STRSIM_API void strsim_b200_cache_clear(void);
STRSIM_API void strsim_b200_cache_stats(int64_t out[4]);
/* Companion measures (arrow_plugin.cpp): the plugin call that uploads a pair of columns also computes -- in
 * the same fused pass, downloads overlapped with the upload -- the measures the PREVIOUS query asked for
 * on its pair of columns, and the later calls of the query (README.md:47-51: five expressions, five calls)
 * take their result ready-made.  Learned from the calls themselves, never guessed; results wait at most
 * the cache's time-to-live.  `enabled` switches the mechanism at run time (initial state: on, unless
 * STRSIM_B200_SPECULATE=0 or the cache is off) and forgets what was learned; returns the previous state.
 * stats: [0] results handed over ready-made, [1] results waiting, [2] bit mask of the learned companions. */
STRSIM_API int strsim_b200_speculation(int enabled);
STRSIM_API void strsim_b200_speculation_stats(int64_t out[3]);
/* thread-local, NUL-terminated description of the last failure on this thread */
STRSIM_API const char *strsim_b200_last_error(void);
/* kernels launched by this library since load (all threads); bench.py reports the delta */
STRSIM_API uint64_t strsim_b200_kernel_launches(void);
/* rows that left the fused short-string kernel for a follow-up kernel in the last call on this
 * thread: [0] 33..64-byte rows, [1] long / generic rows */
STRSIM_API void strsim_b200_last_overflow(int64_t out[2]);
/* row slices of the last host call on this thread that were computed a second time because the
 * progressive upload had not delivered their payload yet (0 for columns laid out sequentially) */
STRSIM_API int strsim_b200_last_redo_slices(void);
STRSIM_API const char *strsim_b200_version(void);

/* ---- Polars plugin ABI (polars-ffi 0.43.1 `version_0`) ------------------------------ */
typedef struct strsim_series_export {
    struct ArrowSchema *field;
    struct ArrowArray **arrays;
    size_t len; /* number of chunks */
    void (*release)(struct strsim_series_export *);
    void *private_data;
} strsim_series_export;

typedef struct strsim_caller_context {
    uint64_t bitflags; /* bit 0: engine is already parallel (strsim.rs:53); irrelevant for a GPU launch */
} strsim_caller_context;

#define STRSIM_DECLARE_PLUGIN(name)                                                                   \
    STRSIM_API void _polars_plugin_##name(strsim_series_export *inputs, size_t n_inputs,              \
                                          const uint8_t *kwargs, size_t kwargs_len,                   \
                                          strsim_series_export *return_value,                         \
                                          strsim_caller_context *context);                            \
    STRSIM_API void _polars_plugin_field_##name(struct ArrowSchema *input_fields, size_t n_fields,    \
                                                struct ArrowSchema *return_field);

STRSIM_DECLARE_PLUGIN(levenshtein)   /* src/expressions/mod.rs:8-11  */
STRSIM_DECLARE_PLUGIN(jaro)          /* src/expressions/mod.rs:13-16 */
STRSIM_DECLARE_PLUGIN(jaro_winkler)  /* src/expressions/mod.rs:18-21 */
STRSIM_DECLARE_PLUGIN(jaccard)       /* src/expressions/mod.rs:23-26 */
STRSIM_DECLARE_PLUGIN(sorensen_dice) /* src/expressions/mod.rs:28-31 */

STRSIM_API const char *_polars_plugin_get_last_error_message(void);
STRSIM_API uint32_t _polars_plugin_get_version(void);

#ifdef __cplusplus
}
#endif
#endif /* STRSIM_B200_H */
