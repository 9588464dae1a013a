// hz_cli.cpp — `datacomp` command line: same verbs, argument order, output lines and exit codes
// as the reference CLI (cli/DataCompCLI.java:24-169), but the codec work runs on the B200 through
// libhuffb200 instead of `new CpuCompressionService(chunkSizeMB)` (:62).
//   datacomp compress|c   <input-file> <output-file> [chunk-size-MB=32]
//   datacomp decompress|d <input-file> <output-file>
// Extension: a chunk size suffixed with 'k' or 'b' (e.g. 64k, 4096b) selects a bytes-granular
// chunk for the 64 KB - 4 MB sweep of BASELINE.json (the container's chunk field is in bytes).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include "huffb200.h"

static void usage() {
    puts("DataComp - GPU-Accelerated Compression Tool");
    puts("");
    puts("Usage:");
    puts("  Compress:   datacomp compress <input-file> <output-file> [chunk-size-MB]");
    puts("  Decompress: datacomp decompress <input-file> <output-file>");
    puts("");
    puts("Examples:");
    puts("  datacomp compress data.tar data.tar.dc");
    puts("  datacomp compress large-file.bin /tmp/compressed.dc 8");
    puts("  datacomp decompress data.tar.dc data-restored.tar");
    puts("");
    puts("Short forms:");
    puts("  'c' for compress, 'd' for decompress");
}

static std::string fmt_size(long long b) {
    char buf[64];
    if (b < 1024) snprintf(buf, sizeof buf, "%lld B", b);
    else if (b < 1024 * 1024) snprintf(buf, sizeof buf, "%.2f KB", b / 1024.0);
    else if (b < 1024LL * 1024 * 1024) snprintf(buf, sizeof buf, "%.2f MB", b / (1024.0 * 1024));
    else snprintf(buf, sizeof buf, "%.2f GB", b / (1024.0 * 1024 * 1024));
    return buf;
}

static long long file_size(const char* p) { struct stat st; return stat(p, &st) == 0 ? (long long)st.st_size : -1; }

static void on_progress(double f, void*) { printf("\rProgress: %d%%", (int)(f * 100)); fflush(stdout); }

int main(int argc, char** argv) {
    if (argc < 4) { usage(); return 1; }
    std::string op = argv[1];
    for (auto& ch : op) ch = (char)tolower(ch);
    const char* in = argv[2];
    const char* out = argv[3];
    unsigned long long chunk_bytes = 32ull * 1024 * 1024;           // default 32 MB (DataCompCLI.java:35)
    if (argc > 4) {
        char* end = nullptr;
        long long v = strtoll(argv[4], &end, 10);
        if (end == argv[4] || v <= 0) { fprintf(stderr, "Invalid chunk size: %s\n", argv[4]); return 1; }
        if (*end == 'k' || *end == 'K') chunk_bytes = (unsigned long long)v * 1024;
        else if (*end == 'b' || *end == 'B') chunk_bytes = (unsigned long long)v;
        else if (*end == 0) chunk_bytes = (unsigned long long)v * 1024 * 1024;
        else { fprintf(stderr, "Invalid chunk size: %s\n", argv[4]); return 1; }
        if (chunk_bytes > 0x7fffffffull) { fprintf(stderr, "Invalid chunk size: %s\n", argv[4]); return 1; }
    }
    if (file_size(in) < 0) { fprintf(stderr, "Error: Input file does not exist: %s\n", in); return 1; }
    if (op != "compress" && op != "c" && op != "decompress" && op != "d") {
        fprintf(stderr, "Unknown operation: %s\n", op.c_str());
        usage();
        return 1;
    }
    hz_ctx* ctx = nullptr;
    int dev = getenv("HZ_DEVICE") ? atoi(getenv("HZ_DEVICE")) : 0;
    if (hz_create(dev, &ctx) != HZ_OK) { fprintf(stderr, "Error: no usable CUDA device (there is no CPU fallback)\n"); return 1; }
    auto t0 = std::chrono::steady_clock::now();
    int rc;
    if (op == "compress" || op == "c") {
        long long isz = file_size(in);
        printf("Compressing...\n  Input:  %s\n  Output: %s\n  Size:   %s\n", in, out, fmt_size(isz).c_str());
        rc = hz_compress_file(ctx, in, out, (uint32_t)chunk_bytes, nullptr, -1, on_progress, nullptr);
        if (rc == HZ_OK) {
            double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            long long osz = file_size(out);
            printf("\n\nCompression complete!\n");
            printf("  Original size:   %s\n", fmt_size(isz).c_str());
            printf("  Compressed size: %s\n", fmt_size(osz).c_str());
            printf("  Compression ratio: %.2f%%\n", isz ? 100.0 * osz / isz : 0.0);
            printf("  Time: %.2f seconds\n", sec);
            printf("  Throughput: %.2f MB/s\n", sec > 0 ? isz / 1e6 / sec : 0.0);
        }
    } else {
        long long isz = file_size(in);
        printf("Decompressing...\n  Input:  %s\n  Output: %s\n", in, out);
        rc = hz_decompress_file(ctx, in, out, on_progress, nullptr);
        if (rc == HZ_OK) {
            double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            long long osz = file_size(out);
            printf("\n\nDecompression complete!\n");
            printf("  Compressed size:   %s\n", fmt_size(isz).c_str());
            printf("  Decompressed size: %s\n", fmt_size(osz).c_str());
            printf("  Time: %.2f seconds\n", sec);
            printf("  Throughput: %.2f MB/s\n", sec > 0 ? osz / 1e6 / sec : 0.0);
        }
    }
    if (rc != HZ_OK) fprintf(stderr, "\nError: %s\n", hz_last_error(ctx));
    hz_destroy(ctx);
    return rc == HZ_OK ? 0 : 1;
}
