package com.datacomp.service.b200;

import com.datacomp.service.FrequencyService;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

/**
 * FrequencyService (service/FrequencyService.java:6-27) backed by the sm_100a histogram kernel of
 * libhuffb200.so.  Same contract as CpuFrequencyService.computeHistogram
 * (service/cpu/CpuFrequencyService.java:29-46): counts of the unsigned byte values of
 * data[offset, offset+length) as long[256].
 *
 * A context is single-owner; the reference calls FrequencyService from up to 8 pool threads
 * (service/cpu/CpuCompressionService.java:42-44), so calls are serialised on the context.
 */
public final class B200FrequencyService implements FrequencyService, AutoCloseable {
    private final MemorySegment ctx;

    public B200FrequencyService(int device) {
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
    public synchronized long[] computeHistogram(byte[] data, int offset, int length) {
        long[] hist = new long[256];
        if (length == 0) return hist;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = a.allocate(length);
            MemorySegment.copy(data, offset, in, JAVA_BYTE, 0, length);
            MemorySegment out = a.allocate(256 * 4L, 4);
            int rc = (int) HuffB200.hz_histogram.invokeExact(ctx, in, (long) length, length, out);   // one chunk
            if (rc != HuffB200.HZ_OK) throw new IllegalStateException(HuffB200.lastError(ctx));
            for (int i = 0; i < 256; i++) hist[i] = Integer.toUnsignedLong(out.getAtIndex(JAVA_INT, i));
            return hist;
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    @Override
    public String getServiceName() {
        return "B200 (CUDA sm_100a)";
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
