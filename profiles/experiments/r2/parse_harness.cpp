#include <chrono>
#include <cstdio>
#include <string>
#include <vector>
#include <cstdint>
#include <cstdlib>
#include "../../../xspect2_b200/csrc/xs_fastx.cpp"
int xs_set_error(int code, const std::string& msg) { fprintf(stderr, "err %d %s\n", code, msg.c_str()); return code; }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    const char* path = argv[1]; unsigned thr = atoi(argv[2]); uint64_t block = 96ull << 20;
    for (int rep = 0; rep < 3; ++rep) {
        xs_fastx* fx = nullptr;
        double t0 = now();
        xs_fastx_open_stream(path, 2, &fx);
        uint64_t fsize = xs_fastx_file_size(fx);
        std::vector<uint8_t> staging(block + (1 << 20));
        std::vector<FastxSegOut> segs;
        uint64_t a = xs_fastx_sync(fx, 0), nrec = 0, nb = 0;
        double tp = 0;
        while (a < fsize) {
            uint64_t b = a + block >= fsize ? fsize : xs_fastx_sync(fx, a + block);
            double t1 = now();
            int rc = xs_fastx_parse_block_1pass(fx, a, b, thr, staging.data(), segs, 21);
            tp += now() - t1;
            if (rc) return 1;
            for (auto& so : segs) { nrec += so.n_rec; nb += so.n_bases; }
            a = b;
        }
        double dt = now() - t0;
        printf("threads %u: %llu records, %llu bases, total %.3f s (parse %.3f) -> %.2f GB/s\n", thr, (unsigned long long)nrec, (unsigned long long)nb, dt, tp, fsize / dt / 1e9);
        xs_fastx_seg_free(segs);
        xs_fastx_close(fx);
    }
}
