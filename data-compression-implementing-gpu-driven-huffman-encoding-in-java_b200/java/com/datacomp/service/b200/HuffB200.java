package com.datacomp.service.b200;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.lang.invoke.MethodHandles;
import java.lang.invoke.MethodType;
import java.nio.file.Path;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (java.lang.foreign, final in JDK 22; preview in the reference's JDK 21 toolchain:
 * compile and run with --enable-preview there) downcall handles for libhuffb200.so — one handle
 * per entry point of include/huffb200.h.  No JNI glue, no generated code: the C ABI takes plain
 * pointers and sizes.
 *
 * The library is located through -Dhuffb200.lib=/path/to/libhuffb200.so (or java.library.path).
 * There is no CPU fallback: when the library or a CUDA device is missing, construction of the
 * services below fails and ServiceFactory keeps the reference's CPU service.
 */
final class HuffB200 {
    static final int HZ_OK = 0;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            Path.of(System.getProperty("huffb200.lib", "libhuffb200.so")), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(
                () -> new UnsatisfiedLinkError("libhuffb200.so does not export " + name)), fd);
    }

    // int hz_create(int device, hz_ctx** out_ctx)
    static final MethodHandle hz_create = h("hz_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    // void hz_destroy(hz_ctx*)
    static final MethodHandle hz_destroy = h("hz_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    // const char* hz_last_error(const hz_ctx*)
    static final MethodHandle hz_last_error = h("hz_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    // int hz_device_count(void)
    static final MethodHandle hz_device_count = h("hz_device_count", FunctionDescriptor.of(JAVA_INT));
    // int hz_histogram(hz_ctx*, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint32_t* hist)
    static final MethodHandle hz_histogram = h("hz_histogram",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS));
    // int hz_build_codebooks(hz_ctx*, const uint32_t* hist, uint32_t K, uint8_t* len, uint32_t* code)
    static final MethodHandle hz_build_codebooks = h("hz_build_codebooks",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    // int hz_encode(hz_ctx*, in, n, chunk_bytes, out, out_cap, comp_off, len_out, hist_out)
    static final MethodHandle hz_encode = h("hz_encode", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG,
            JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    // int hz_decode(hz_ctx*, comp, comp_bytes, comp_off, comp_size, orig_size, orig_off, len, K, out, out_cap)
    static final MethodHandle hz_decode = h("hz_decode", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG));
    // int hz_sha256_chunks(hz_ctx*, in, n, chunk_bytes, digests)
    static final MethodHandle hz_sha256_chunks = h("hz_sha256_chunks",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS));
    // int hz_compress_file(hz_ctx*, in_path, out_path, chunk_bytes, name_override, mtime_ms_override, progress, user)
    static final MethodHandle hz_compress_file = h("hz_compress_file", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS,
            ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    // int hz_decompress_file(hz_ctx*, in_path, out_path, progress, user)
    static final MethodHandle hz_decompress_file = h("hz_decompress_file",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    // int hz_verify_file(hz_ctx*, path, int* ok)
    static final MethodHandle hz_verify_file = h("hz_verify_file",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));

    // int hz_stage_metrics(const hz_ctx*, hz_stage_metric out[8])   {double ms; uint64_t count; uint64_t bytes;}
    static final MethodHandle hz_stage_metrics = h("hz_stage_metrics", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    // int hz_encode_global(hz_ctx*, in, n, chunk_bytes, out, out_cap, comp_off, len256_out)   (multi-GPU extension mode)
    static final MethodHandle hz_encode_global = h("hz_encode_global", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS,
            JAVA_LONG, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    // int hz_comm_unique_id(void* id128); int hz_comm_init(hz_ctx*, const void* id128, int nranks, int rank)
    static final MethodHandle hz_comm_unique_id = h("hz_comm_unique_id", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle hz_comm_init = h("hz_comm_init", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));

    /** typedef void (*hz_progress_fn)(double fraction, void* user) */
    static final FunctionDescriptor PROGRESS_FN = FunctionDescriptor.ofVoid(JAVA_DOUBLE, ADDRESS);

    /** Upcall stub that forwards the library's per-chunk progress to a Java callback. */
    static MemorySegment progressStub(java.util.function.Consumer<Double> cb, Arena arena) {
        if (cb == null) return MemorySegment.NULL;
        try {
            MethodHandle target = MethodHandles.lookup().findStatic(HuffB200.class, "onProgress",
                    MethodType.methodType(void.class, java.util.function.Consumer.class, double.class, MemorySegment.class));
            return LINKER.upcallStub(target.bindTo(cb), PROGRESS_FN, arena);
        } catch (ReflectiveOperationException e) {
            throw new IllegalStateException(e);
        }
    }

    @SuppressWarnings("unused")
    private static void onProgress(java.util.function.Consumer<Double> cb, double fraction, MemorySegment user) {
        cb.accept(fraction);
    }

    static String lastError(MemorySegment ctx) {
        try {
            MemorySegment p = (MemorySegment) hz_last_error.invokeExact(ctx);
            return p.reinterpret(512).getString(0);      // getUtf8String(0) on the JDK 21 preview API
        } catch (Throwable t) {
            return "unknown error";
        }
    }

    private HuffB200() { }
}
