// Multi-GPU demo of the C ABI from a C++ host: one thread, one r360_ctx and one NCCL rank per GPU; every rank
// registers its own shard of synthetic pairs and all ranks end up with all result records.
//   g++ -std=c++17 -I include tests/cpp/allgather_demo.cpp -o allgather_demo \
//       rgbd360_b200/librgbd360_b200.so -lnccl -L/usr/local/cuda/lib64 -lcudart -lpthread
//   ./allgather_demo [n_gpus = 2] [pairs_per_gpu = 4]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <nccl.h>
#include "r360.h"

static int g_fail = 0;
#define REQUIRE(cond, ...) do { if (!(cond)) { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); ++g_fail; return; } } while (0)

static void rank_main(int rank, int n_ranks, int n_local, ncclComm_t comm, std::vector<r360_result>* all_out) {
    const int rows = 128, cols = 256, n_frames = 2 * n_local;
    r360_params P;
    r360_default_params(&P);
    P.n_levels = 3;
    r360_ctx* ctx = nullptr;
    REQUIRE(r360_create(&ctx, rank, rows, cols, n_frames, n_local, &P) == R360_OK, "rank %d: create: %s", rank, r360_last_error(nullptr));
    // pair p of this rank = (target frame 2g, source frame 2g + 1), g = rank * n_local + p: every GPU registers other pairs
    std::vector<uint8_t> rgb((size_t)n_frames * rows * cols * 3);
    std::vector<uint16_t> depth((size_t)n_frames * rows * cols);
    REQUIRE(r360_synth_frames(ctx, 0, 2 * rank * n_local, n_frames, rgb.data(), depth.data()) == R360_OK, "rank %d: synth: %s", rank, r360_last_error(ctx));
    REQUIRE(r360_set_frames(ctx, 0, n_frames, rgb.data(), depth.data(), nullptr) == R360_OK, "rank %d: set_frames: %s", rank, r360_last_error(ctx));
    std::vector<int32_t> src(n_local), trg(n_local);
    for (int p = 0; p < n_local; ++p) { trg[p] = 2 * p; src[p] = 2 * p + 1; }
    std::vector<r360_result> local(n_local);
    REQUIRE(r360_register_pairs(ctx, n_local, src.data(), trg.data(), nullptr, local.data(), nullptr) == R360_OK, "rank %d: register: %s", rank, r360_last_error(ctx));
    for (int p = 0; p < n_local; ++p) local[p].pair_id = rank * n_local + p;          // global ids
    all_out->assign((size_t)n_ranks * n_local, r360_result());
    REQUIRE(r360_allgather_results(ctx, comm, local.data(), n_local, n_ranks, all_out->data()) == R360_OK, "rank %d: allgather: %s", rank, r360_last_error(ctx));
    // this rank's own records must come back where they belong, bit for bit
    for (int p = 0; p < n_local; ++p) {
        const r360_result& a = (*all_out)[(size_t)rank * n_local + p];
        for (int k = 0; k < 16; ++k) REQUIRE(a.pose[k] == local[p].pose[k], "rank %d: own record %d changed in the gather", rank, p);
    }
    r360_destroy(ctx);
}

int main(int argc, char** argv) {
    const int n_ranks = argc > 1 ? std::atoi(argv[1]) : 2, n_local = argc > 2 ? std::atoi(argv[2]) : 4;
    std::vector<ncclComm_t> comms(n_ranks);
    std::vector<int> devs(n_ranks);
    for (int r = 0; r < n_ranks; ++r) devs[r] = r;
    if (ncclCommInitAll(comms.data(), n_ranks, devs.data()) != ncclSuccess) { std::fprintf(stderr, "ncclCommInitAll failed (needs %d GPUs)\n", n_ranks); return 2; }
    std::vector<std::vector<r360_result>> all(n_ranks);
    std::vector<std::thread> th;
    for (int r = 0; r < n_ranks; ++r) th.emplace_back(rank_main, r, n_ranks, n_local, comms[r], &all[r]);
    for (auto& t : th) t.join();
    if (g_fail) return 1;
    // every rank holds the same complete list: ids 0 .. n-1 in order, all registered, identical bits everywhere
    const int n = n_ranks * n_local;
    for (int r = 0; r < n_ranks; ++r)
        for (int q = 0; q < n; ++q) {
            const r360_result& a = all[r][q];
            const r360_result& b = all[0][q];
            if (a.pair_id != q || a.status != R360_PAIR_OK) { std::fprintf(stderr, "rank %d record %d: pair_id %d status %d\n", r, q, a.pair_id, a.status); return 1; }
            for (int k = 0; k < 16; ++k) if (a.pose[k] != b.pose[k]) { std::fprintf(stderr, "rank %d record %d differs from rank 0\n", r, q); return 1; }
            double T[16];
            r360_synth_gt_pose(0, 2 * q + 1, 2 * q, T);                                 // analytic ground truth T_trg<-src
            double dt = 0;
            for (int k = 12; k < 15; ++k) dt = std::fmax(dt, std::fabs(T[k] - (double)a.pose[k]));
            if (dt > 0.02) { std::fprintf(stderr, "record %d: translation off the ground truth by %.4f m\n", q, dt); return 1; }
        }
    for (auto c : comms) ncclCommDestroy(c);
    std::printf("allgather_demo ok: %d ranks x %d pairs, every rank holds all %d records\n", n_ranks, n_local, n);
    return 0;
}
