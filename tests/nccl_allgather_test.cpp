// nccl_allgather_test.cpp — a C++ host running the multi-GPU exchange step through the C ABI alone
// (include/slamb200.h: sb_nccl_unique_id / sb_nccl_comm_init / sb_allgather_kf_poses), no MPI, no Python:
// one thread per GPU of this box, round-robin keyframe ownership, every rank must end with the full pose table.
//   usage: nccl_allgather_test <n_gpus> <n_keyframes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/slamb200.h"

static double pose_value(int kf, int c) { return kf * 10.0 + c + 0.25; }

int main(int argc, char **argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 2, n_kf = argc > 2 ? atoi(argv[2]) : 742;
    int version = 0;
    if (sb_nccl_version(&version) != SB_OK) { printf("FAIL sb_nccl_version: %s\n", sb_last_error()); return 1; }
    uint8_t id[128];
    if (sb_nccl_unique_id(id) != SB_OK) { printf("FAIL sb_nccl_unique_id: %s\n", sb_last_error()); return 1; }
    const int cap = (n_kf + world - 1) / world;
    std::vector<int> status(world, -100);
    std::vector<std::thread> threads;
    for (int rank = 0; rank < world; ++rank)
        threads.emplace_back([&, rank]() {
            void *comm = nullptr;
            int rc = sb_nccl_comm_init(&comm, rank, world, rank, id);
            if (rc != SB_OK) { printf("rank %d: sb_nccl_comm_init: %s\n", rank, sb_last_error()); status[rank] = rc; return; }
            std::vector<double> local((size_t)cap * 7, 0.0), all((size_t)world * cap * 7, -1.0);
            std::vector<int> counts(world, -1);
            int n_local = 0;
            for (int kf = rank; kf < n_kf; kf += world, ++n_local)
                for (int c = 0; c < 7; ++c) local[(size_t)n_local * 7 + c] = pose_value(kf, c);
            rc = sb_allgather_kf_poses(comm, nullptr, local.data(), n_local, all.data(), counts.data(), cap);
            if (rc != SB_OK) { printf("rank %d: sb_allgather_kf_poses: %s\n", rank, sb_last_error()); status[rank] = rc; return; }
            int bad = 0;
            for (int r = 0; r < world; ++r) bad += counts[r] != (n_kf - r + world - 1) / world;
            for (int kf = 0; kf < n_kf; ++kf)
                for (int c = 0; c < 7; ++c) bad += all[((size_t)(kf % world) * cap + kf / world) * 7 + c] != pose_value(kf, c);
            // argument checks come back as status codes, never as exceptions
            bad += sb_allgather_kf_poses(comm, nullptr, local.data(), cap + 1, all.data(), counts.data(), cap) != SB_ERR_INVALID;
            bad += sb_allgather_kf_poses(nullptr, nullptr, local.data(), 0, all.data(), counts.data(), cap) != SB_ERR_INVALID;
            sb_nccl_comm_destroy(comm);
            status[rank] = bad;
        });
    for (auto &t : threads) t.join();
    int bad = 0;
    for (int r = 0; r < world; ++r) bad += status[r] != 0;
    printf("%s: NCCL %d, %d ranks, %d keyframes, cap %d, status", bad ? "FAIL" : "OK", version, world, n_kf, cap);
    for (int r = 0; r < world; ++r) printf(" %d", status[r]);
    printf("\n");
    return bad ? 1 : 0;
}
