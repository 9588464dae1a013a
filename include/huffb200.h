/* huffb200.h — C ABI of the B200-native Huffman codec (libhuffb200.so).
 *
 * This is the drop-in boundary for the ONE hot path of DataComp
 * (vuyraj/Data-Compression-Implementing-GPU-Driven-Huffman-Encoding-in-Java):
 *   byte histogram -> canonical Huffman codebook -> MSB-first bit-packed encode -> chunked decode,
 * batched over all chunks of a file and executed by hand-written sm_100a CUDA kernels.
 * Plain pointers and sizes only; no torch / C++ types.  A Java host binds these with Panama FFM
 * (Linker.downcallHandle) or JNI — see INTEGRATION.md for the stubs.
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * app/src/main/java/com/datacomp/ in the reference tree).
 *
 * Pointer kinds: every data pointer may be HOST or DEVICE memory (detected with
 * cudaPointerGetAttributes).  Device pointers are used in place; host pointers are staged
 * through context-owned device buffers.  All work is enqueued on the context's stream.  A call
 * whose outputs are all device memory returns without synchronising (errors detected on the
 * device are then reported by hz_sync / the next synchronising call); a call with any host
 * output synchronises before returning.
 *
 * There is NO CPU fallback: if no CUDA device is usable, hz_create fails.
 */
#ifndef HUFFB200_H
#define HUFFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hz_ctx hz_ctx;

/* status codes (0 = ok, negatives = errors) */
#define HZ_OK 0
#define HZ_ERR_ARG (-1)           /* bad argument                                                  */
#define HZ_ERR_CUDA (-2)          /* CUDA runtime error (text in hz_last_error)                    */
#define HZ_ERR_NOMEM (-3)         /* device or host allocation failed                              */
#define HZ_ERR_CODE_TOO_LONG (-4) /* a chunk needs a codeword > 32 bits; the reference throws
                                     ArrayIndexOutOfBounds at core/CanonicalHuffman.java:107       */
#define HZ_ERR_OUT_TOO_SMALL (-5) /* output capacity too small                                     */
#define HZ_ERR_DECODE (-6)        /* bit pattern matches no codeword: "Huffman decode error at
                                     position i", core/TableBasedHuffmanDecoder.java:109-111       */
#define HZ_ERR_BAD_LENGTHS (-7)   /* code-length table is not a prefix code / length > 32          */
#define HZ_ERR_IO (-8)            /* file I/O failed                                               */
#define HZ_ERR_FORMAT (-9)        /* bad magic / version / footer pointer,
                                     core/CompressionHeader.java:93-99, cpu/CpuCompressionService.java:372-374 */
#define HZ_ERR_CHECKSUM (-10)     /* per-chunk SHA-256 mismatch, cpu/CpuCompressionService.java:537-550 */
#define HZ_ERR_UNSUPPORTED (-11)  /* resumeCompression: cpu/CpuCompressionService.java:636-641 throws too */

/* ---- context ------------------------------------------------------------------------------ */

/* Create a codec context on CUDA device `device` (one context per GPU; single-owner, i.e. one
 * host thread at a time).  Replaces the constructor of the service objects
 * (cpu/CpuCompressionService.java:36-47, gpu/GpuCompressionService.java ctor).               */
int hz_create(int device, hz_ctx** out_ctx);
void hz_destroy(hz_ctx* ctx);                       /* AutoCloseable.close(), cpu/...:771-790     */
const char* hz_last_error(const hz_ctx* ctx);       /* message of the last failing call           */
const char* hz_strerror(int status);
int hz_set_stream(hz_ctx* ctx, void* cuda_stream);  /* cudaStream_t; NULL = context's own stream  */
int hz_sync(hz_ctx* ctx);                           /* wait for enqueued work, return device-side status */
int hz_device_count(void);                          /* isAvailable(): service/CompressionService.java:65 */
uint32_t hz_version(void);

/* number of chunks: (n + chunk - 1) / chunk, cpu/CpuCompressionService.java:64 */
uint64_t hz_num_chunks(uint64_t n, uint32_t chunk_bytes);

/* ---- stage level (one call = all chunks of a buffer) -------------------------------------- */

/* Per-chunk byte histogram.  hist is K x 256 uint32 (a chunk is < 2^32 bytes).
 * Replaces FrequencyService.computeHistogram(byte[],int,int) -> long[256]
 * (service/FrequencyService.java:16; cpu/CpuFrequencyService.java:29-46;
 *  gpu/GpuFrequencyService.java:87-149), once per chunk.                                       */
int hz_histogram(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint32_t* hist);

/* Canonical Huffman codebooks from K histograms: len[K][256] (0 = absent symbol, 1..32) and
 * code[K][256] (right-aligned codewords; bit len-1 is emitted first).  Bit-identical to
 * CanonicalHuffman.buildCanonicalCodes(long[256]) (core/CanonicalHuffman.java:19-132), including
 * the java.util.PriorityQueue tie-breaking.  `code` may be NULL.                               */
int hz_build_codebooks(hz_ctx* ctx, const uint32_t* hist, uint32_t K, uint8_t* len, uint32_t* code);

/* Codewords from stored lengths: CanonicalHuffman.generateCanonicalCodesFromLengths(int[256])
 * (core/CanonicalHuffman.java:141-146).                                                        */
int hz_codes_from_lengths(hz_ctx* ctx, const uint8_t* len, uint32_t K, uint32_t* code);

/* Full per-chunk pipeline histogram -> codebook -> encode for all K = hz_num_chunks(n, chunk_bytes)
 * chunks of `in`.  `out` receives the dense payload section of a .dcz file: chunk k's
 * zero-padded MSB-first bitstream at out[comp_off[k] .. comp_off[k+1]).  comp_off has K+1
 * entries (comp_off[K] = total payload bytes).  len_out is K x 256 (the footer's code lengths);
 * hist_out (K x 256 uint32) may be NULL.  out_cap >= n always suffices (mean Huffman length <= 8).
 * Replaces CpuCompressionService.processChunk minus SHA-256 (cpu/CpuCompressionService.java:
 * 233-260: computeHistogram, buildCanonicalCodes, encodeChunk :303-315 / BitOutputStream :711-737)
 * and GpuCompressionService.processChunkGpu (gpu/GpuCompressionService.java:389-460).           */
int hz_encode(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes,
              uint8_t* out, uint64_t out_cap, uint64_t* comp_off, uint8_t* len_out, uint32_t* hist_out);

/* Same, but every chunk is coded with ONE caller-supplied length table len256[256] (global-codebook
 * extension for multi-GPU runs: the caller all-reduces the histograms, builds one codebook with
 * hz_build_codebooks and passes its lengths here).  Not bit-identical to the reference
 * compressor, but a valid .dcz any reference decoder accepts.                                   */
int hz_encode_with_lengths(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes,
                           const uint8_t* len256, uint8_t* out, uint64_t out_cap, uint64_t* comp_off);

/* Global-codebook extension for one logical file sharded over several GPUs (one context per GPU, one process
 * per GPU): ONE codebook, built from the byte histogram summed over ALL chunks of ALL ranks, codes every chunk.
 * hz_comm_unique_id (on one rank; the host distributes the 128 bytes) + hz_comm_init (on every rank) create an
 * NCCL communicator inside the library (libnccl.so.2 is bound with dlopen at that moment; without these calls
 * the sum is over this context's chunks only).  hz_encode_global then runs histogram -> device-side reduction to
 * 256 x u64 -> ncclAllReduce(sum) over NVLink on the context's stream -> the same deterministic codebook build on
 * every rank -> encode, all enqueued without host synchronisation; len256_out receives the 256 code lengths
 * (identical on every rank; the footer repeats them in every chunk record).  Counts beyond 32 bits are scaled by a
 * power of two first.  Collective: every rank must call it, also with n == 0.  Not bit-identical to the reference
 * compressor (which has no such mode), but a valid .dcz for any reference decoder.                             */
int hz_comm_unique_id(void* id128);
int hz_comm_init(hz_ctx* ctx, const void* id128, int nranks, int rank);
int hz_comm_destroy(hz_ctx* ctx);
int hz_encode_global(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint8_t* out, uint64_t out_cap,
                     uint64_t* comp_off, uint8_t* len256_out);

/* Decode K chunks.  Chunk k's bitstream is comp[comp_off[k] .. comp_off[k]+comp_size[k]), its
 * code lengths len[k][256]; exactly orig_size[k] symbols are produced at out[orig_off[k] ..)
 * (orig_off == NULL: chunks are written back to back).  Bits past the end of a chunk read as 0
 * (core/TableBasedHuffmanDecoder.java:204-208).  comp_bytes = bytes addressable from `comp`.
 * Replaces CpuCompressionService.decodeChunkParallel's decode (cpu/CpuCompressionService.java:
 * 511-532): generateCanonicalCodesFromLengths + TableBasedHuffmanDecoder.decode
 * (core/TableBasedHuffmanDecoder.java:103-152).                                                */
int hz_decode(hz_ctx* ctx, const uint8_t* comp, uint64_t comp_bytes, const uint64_t* comp_off,
              const uint32_t* comp_size, const uint32_t* orig_size, const uint64_t* orig_off,
              const uint8_t* len, uint32_t K, uint8_t* out, uint64_t out_cap);

/* Page-locked host memory.  hz_encode / hz_decode called with HOST input and output of >= 128 MiB
 * pipeline the transfers in batches of whole chunks (H2D of batch b+1, kernels of batch b and D2H
 * of batch b-1 overlap on three streams); the copies are only asynchronous from page-locked
 * memory.  A Java host wraps the returned address in a MemorySegment (replaces the byte[] chunk
 * buffers of cpu/CpuCompressionService.java:214-224).                                            */
int hz_host_alloc(void** p, size_t bytes);
void hz_host_free(void* p);

/* SHA-256 of each chunk: digests is K x 32 bytes.  Replaces ChecksumUtil.computeSha256 per chunk
 * (util/ChecksumUtil.java:11-27; cpu/CpuCompressionService.java:226-228, :536).                 */
int hz_sha256_chunks(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint8_t* digests);

/* ---- file level (the CompressionService surface) ------------------------------------------- */

typedef void (*hz_progress_fn)(double fraction, void* user);   /* Consumer<Double> progressCallback */

/* CompressionService.compress(Path, Path, Consumer<Double>) (service/CompressionService.java:21-22;
 * cpu/CpuCompressionService.java:57-205): writes a byte-identical .dcz (payload, footer, 8-byte
 * footer pointer).  chunk_bytes = chunkSizeMB*1024*1024 for the reference constructor.
 * name_override / mtime_ms_override (NULL / <0 = take from the input file) exist for tests.      */
int hz_compress_file(hz_ctx* ctx, const char* in_path, const char* out_path, uint32_t chunk_bytes,
                     const char* name_override, int64_t mtime_ms_override,
                     hz_progress_fn progress, void* user);

/* CompressionService.decompress (service/CompressionService.java:33-34;
 * cpu/CpuCompressionService.java:318-506): header-first probe then footer-last, decode on the GPU,
 * verify every chunk's SHA-256.                                                                 */
int hz_decompress_file(hz_ctx* ctx, const char* in_path, const char* out_path,
                       hz_progress_fn progress, void* user);

/* CompressionService.verifyIntegrity (service/CompressionService.java:55): decodes every chunk on
 * the GPU and checks its SHA-256 without writing output.  *ok = 1/0.                            */
int hz_verify_file(hz_ctx* ctx, const char* path, int* ok);

/* In-memory variants of the two above (host buffers).  *out is malloc'd; free with hz_free.    */
int hz_compress_buffer(hz_ctx* ctx, const uint8_t* data, uint64_t n, uint32_t chunk_bytes,
                       const char* name, int64_t mtime_ms, uint8_t** out, uint64_t* out_n);
int hz_decompress_buffer(hz_ctx* ctx, const uint8_t* dcz, uint64_t n, uint8_t** out, uint64_t* out_n);
void hz_free(void* p);

/* Stage timings of the LAST file- or buffer-level call (hz_compress_file, hz_decompress_file, hz_verify_file,
 * hz_compress_buffer, hz_decompress_buffer), one entry per model/StageMetrics.Stage constant in declaration order
 * (model/StageMetrics.java:11-20).  Replaces getLastStageMetrics() of the service classes
 * (cpu/CpuCompressionService.java:52, reached from ui/CompressController.java:292-297): the Java shim feeds every
 * entry to StageMetrics.recordStage(stage, ns, bytes).  Kernel stages are CUDA-event times of the kernels that do
 * that stage's work (histogram; codebook build and decode-table rebuild; bit packing; decode); checksum and file
 * I/O are host wall-clock times of the pipeline's tasks (they overlap with the GPU stages).                  */
#define HZ_STAGE_FREQUENCY_ANALYSIS 0
#define HZ_STAGE_HUFFMAN_TREE_BUILD 1
#define HZ_STAGE_ENCODING 2
#define HZ_STAGE_CHECKSUM_COMPUTE 3
#define HZ_STAGE_FILE_IO 4
#define HZ_STAGE_HEADER_WRITE 5
#define HZ_STAGE_DECODING 6
#define HZ_STAGE_CHECKSUM_VERIFY 7
#define HZ_STAGE_COUNT 8
typedef struct hz_stage_metric { double ms; uint64_t count; uint64_t bytes; } hz_stage_metric;
int hz_stage_metrics(const hz_ctx* ctx, hz_stage_metric out[HZ_STAGE_COUNT]);
const char* hz_stage_name(int stage);               /* "FREQUENCY_ANALYSIS", ... (the enum constant's name) */

/* ---- introspection for the bench harness --------------------------------------------------- */

/* Names and accumulated CUDA-event milliseconds / launch counts of the library's kernels since
 * the last hz_prof_reset (only collected while profiling is enabled with hz_prof_enable).       */
int hz_prof_enable(hz_ctx* ctx, int on);
int hz_prof_reset(hz_ctx* ctx);
int hz_prof_count(hz_ctx* ctx);
int hz_prof_get(hz_ctx* ctx, int i, const char** name, double* total_ms, uint64_t* launches);
uint64_t hz_launch_count(const hz_ctx* ctx);       /* kernels launched by this context so far     */

#ifdef __cplusplus
}
#endif
#endif /* HUFFB200_H */
