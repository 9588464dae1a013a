package com.datacomp.service.b200;

import com.datacomp.model.StageMetrics;
import com.datacomp.service.CompressionService;

import java.io.IOException;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.file.Path;
import java.util.function.Consumer;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * CompressionService (service/CompressionService.java:11-66) backed by libhuffb200.so: the per-chunk
 * hot path (histogram, canonical codebook, bit-packed encode, chunked decode) runs on the B200; the
 * per-chunk SHA-256 of the file calls is computed by the library on the HOST's SHA units while the GPU
 * codes the same batch (a handful of 16-32 MiB chunks per batch hash faster there than on the GPU, whose
 * kernel, hz_sha256_chunks, only wins with hundreds of chunks per batch); the .dcz written is
 * byte-identical to CpuCompressionService's
 * (service/cpu/CpuCompressionService.java:57-205) for the same input, chunk size, file name and
 * modification time.  Constructor mirrors CpuCompressionService(int chunkSizeMB) (:36-47);
 * withChunkBytes() adds the bytes-granular chunk size of the 64 KiB - 4 MiB sweeps.
 *
 * Wiring (INTEGRATION.md): ServiceFactory.createCompressionService returns this class when
 * isAvailable(); DataCompCLI constructs it in place of CpuCompressionService.  There is no CPU
 * fallback inside this class.
 */
public final class B200CompressionService implements CompressionService, AutoCloseable {
    private final MemorySegment ctx;
    private final int chunkBytes;

    public B200CompressionService(int chunkSizeMB) {
        this(0, chunkSizeMB * 1024 * 1024);
    }

    public static B200CompressionService withChunkBytes(int device, int chunkBytes) {
        return new B200CompressionService(device, chunkBytes);
    }

    private B200CompressionService(int device, int chunkBytes) {
        if (chunkBytes <= 0) throw new IllegalArgumentException("chunk size must be positive");
        this.chunkBytes = chunkBytes;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) HuffB200.hz_create.invokeExact(device, out);
            if (rc != HuffB200.HZ_OK) throw new IllegalStateException("hz_create failed (" + rc + "): no usable CUDA device");
            ctx = out.get(ADDRESS, 0);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    @Override
    public synchronized void compress(Path inputPath, Path outputPath, Consumer<Double> progressCallback) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            int rc = (int) HuffB200.hz_compress_file.invokeExact(ctx, a.allocateFrom(inputPath.toString()),
                    a.allocateFrom(outputPath.toString()), chunkBytes, MemorySegment.NULL, -1L,
                    HuffB200.progressStub(progressCallback, a), MemorySegment.NULL);
            if (rc != HuffB200.HZ_OK) throw new IOException(HuffB200.lastError(ctx));
        } catch (IOException | RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
    }

    @Override
    public synchronized void decompress(Path inputPath, Path outputPath, Consumer<Double> progressCallback) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            int rc = (int) HuffB200.hz_decompress_file.invokeExact(ctx, a.allocateFrom(inputPath.toString()),
                    a.allocateFrom(outputPath.toString()), HuffB200.progressStub(progressCallback, a), MemorySegment.NULL);
            if (rc != HuffB200.HZ_OK) throw new IOException(HuffB200.lastError(ctx));   // bad magic, checksum mismatch, decode error
        } catch (IOException | RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
    }

    @Override
    public void resumeCompression(Path inputPath, Path outputPath, int lastCompletedChunk,
                                  Consumer<Double> progressCallback) throws IOException {
        throw new UnsupportedOperationException("Resume not yet implemented");   // as cpu/CpuCompressionService.java:636-641
    }

    @Override
    public synchronized boolean verifyIntegrity(Path compressedPath) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment ok = a.allocate(JAVA_INT);
            int rc = (int) HuffB200.hz_verify_file.invokeExact(ctx, a.allocateFrom(compressedPath.toString()), ok);
            if (rc != HuffB200.HZ_OK) throw new IOException(HuffB200.lastError(ctx));
            return ok.get(JAVA_INT, 0) != 0;
        } catch (IOException | RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
    }

    /**
     * Stage timings of the last compress / decompress / verifyIntegrity call, as the reference's services expose
     * them (cpu/CpuCompressionService.java:52; the GUI reaches it by an instanceof cast,
     * ui/CompressController.java:292-297): hz_stage_metrics fills one {double ms, uint64 count, uint64 bytes}
     * per StageMetrics.Stage constant, in declaration order (model/StageMetrics.java:11-20).
     */
    public synchronized StageMetrics getLastStageMetrics() {
        StageMetrics m = new StageMetrics();
        try (Arena a = Arena.ofConfined()) {
            StageMetrics.Stage[] stages = StageMetrics.Stage.values();
            MemorySegment buf = a.allocate(24L * stages.length, 8);
            int rc = (int) HuffB200.hz_stage_metrics.invokeExact(ctx, buf);
            if (rc != HuffB200.HZ_OK) return m;
            for (int i = 0; i < stages.length; i++) {
                double ms = buf.get(JAVA_DOUBLE, 24L * i);
                long count = buf.get(JAVA_LONG, 24L * i + 8), bytes = buf.get(JAVA_LONG, 24L * i + 16);
                if (count > 0) m.recordStage(stages[i], (long) (ms * 1e6), bytes);
            }
        } catch (Throwable t) {
            // metrics are best effort, like the reference's
        }
        return m;
    }

    @Override
    public String getServiceName() {
        return "B200 Compression (CUDA sm_100a)";
    }

    @Override
    public boolean isAvailable() {
        try {
            return (int) HuffB200.hz_device_count.invokeExact() > 0;
        } catch (Throwable t) {
            return false;
        }
    }

    @Override
    public synchronized void close() {
        try {
            HuffB200.hz_destroy.invokeExact(ctx);
        } catch (Throwable ignored) {
        }
    }
}
