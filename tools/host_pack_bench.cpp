// Development aid: host-only timing of the one-pass FASTQ walker and the transfer-form packers (no GPU work).
//   nvcc -O3 -std=c++17 -Xcompiler -pthread -I include -o /tmp/host_pack_bench tools/host_pack_bench.cpp \
//        -L ngs-barcode-count_b200/lib -lbc_b200 -lz -Xlinker -rpath -Xlinker $PWD/ngs-barcode-count_b200/lib
//   /tmp/host_pack_bench reads.fastq THREADS MODE [CHUNK_BYTES]     MODE 0: frame + pack, 1: frame only, 2: frame + bases only,
//                                                                   3: frame + pack into a cache-resident ring (no output stream)
#include "../ngs-barcode-count_b200/csrc/host/bc_host.cpp"
#include <chrono>
int main(int argc, char** argv) {
    const char* path = argv[1];
    unsigned threads = argc > 2 ? atoi(argv[2]) : 8;
    int mode = argc > 3 ? atoi(argv[3]) : 0;  // 0: split+pack, 1: split only, 2: split + planes only
    MappedFile mf;
    if (!mf.open_plain(path)) return 1;
    Pool pool(threads);
    const uint32_t mrl = 150, batch = 1u << 20;
    WireLayout L(batch, mrl, true);
    unsigned char* arena = (unsigned char*)aligned_alloc(4096, (L.total + 4095) & ~4095ul);
    memset(arena, 0, L.total);
    for (int rep = 0; rep < 4; rep++) {
        FusedWalk walk(pool, mf.data, mf.size, argc > 4 ? atoi(argv[4]) : (256u << 10));
        auto t0 = std::chrono::steady_clock::now();
        size_t total = 0;
        std::atomic<uint64_t> sink{0};
        while (walk.more()) {
            total += walk.batch(batch, [&](size_t, const ReadRef* r, size_t base, size_t take) {
                if (mode == 1) { uint64_t s = 0; for (size_t i = 0; i < take; i++) s += r[i].len; sink += s; return; }
                if (mode == 3) { pack_range_wire(mrl, L, arena, r, 0, std::min<size_t>(take, 1), 6); for (size_t i = 0; i < take; i++) pack_range_wire(mrl, L, arena, r + i, i & 63, 1, 6); return; }
                pack_range_wire(mrl, L, arena, r, base, take, mode == 2 ? 0 : 6);
            });
        }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("mode %d threads %u: %zu reads in %.1f ms = %.1f M reads/s, %.0f ns/read/thread\n", mode, threads, total, dt * 1e3, total / dt / 1e6, dt * 1e9 * threads / total);
    }
    return 0;
}
