// CPU check of the file rendezvous the C++ multi-GPU host layer uses between plain processes of one box (FileComm,
// DistEngine.cpp): N processes, rounds of allgather with changing sizes and contents, barriers in between; every rank
// verifies every peer's contribution of every round.  No GPU.
//   for r in 0 1 2; do MIMSEM_RANK=$r MIMSEM_WORLD=3 build/filecomm_check /tmp/rdv & done; wait
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "DistEngine.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const int rank = std::atoi(getenv("MIMSEM_RANK") ? getenv("MIMSEM_RANK") : "0");
    const int world = std::atoi(getenv("MIMSEM_WORLD") ? getenv("MIMSEM_WORLD") : "1");
    int bad = 0;
    try {
        mimsem_host::FileComm comm(argv[1], rank, world);
        for (int round = 0; round < 40; round++) {
            const int64_t n = 1 + (round * 37) % 5000;   // bytes per rank change from round to round
            std::vector<unsigned char> mine(n), all((size_t)n * world);
            for (int64_t i = 0; i < n; i++) mine[i] = (unsigned char)((i * 7 + rank * 13 + round) & 0xff);
            comm.allgather(mine.data(), n, all.data());
            for (int q = 0; q < world; q++)
                for (int64_t i = 0; i < n; i++)
                    if (all[(size_t)q * n + i] != (unsigned char)((i * 7 + q * 13 + round) & 0xff)) bad++;
            if (round % 3 == 0) comm.barrier();
        }
        comm.barrier();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rank %d: %s\n", rank, e.what());
        return 1;
    }
    std::printf("filecomm_check rank %d of %d %s\n", rank, world, bad ? "FAIL" : "ok");
    return bad ? 1 : 0;
}
