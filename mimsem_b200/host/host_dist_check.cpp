// Multi-GPU check of the C++ host layer (DistEngine): N plain processes of one box, rank r on GPU r, rendezvous through
// files.  Every rank applies the partitioned operators to its element block with zeroed ghost rows (so the ghost
// refresh is exercised); rank 0 gathers the owned rows and compares them BITWISE with a one-GPU apply on the whole mesh.
//   for r in 0 1; do MIMSEM_RANK=$r MIMSEM_WORLD=2 build/host_dist_check sphere 3 6 30 /tmp/rdv & done; wait
// A sixth argument "time" adds a timing of the M1 apply (stream order and as bursts of 16 launches) on that mesh and prints
// one JSON line (for the benchmark shape: sphere 4 48 60).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "DistEngine.h"

using namespace mimsem_host;

namespace {

struct SelfComm : Comm {
    SelfComm() { rank = 0; world = 1; }
    void allgather(const void* s, int64_t b, void* r) { std::memcpy(r, s, (size_t)b); }
    void barrier() {}
};

std::vector<double> pseudo_random(size_t n, unsigned long long seed, double lo, double hi) {
    std::vector<double> v(n);
    unsigned long long s = seed * 6364136223846793005ull + 1442695040888963407ull;
    for (size_t i = 0; i < n; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        v[i] = lo + (hi - lo) * (double)(s >> 11) / 9007199254740992.0;
    }
    return v;
}

// rank r contributes its owned rows (zeros elsewhere); the sum over ranks is the global field
void gather_sum(Comm& comm, std::vector<double>& v) {
    std::vector<double> all((size_t)comm.world * v.size());
    comm.allgather(v.data(), (int64_t)v.size() * 8, all.data());
    for (size_t i = 0; i < v.size(); i++) {
        double s = 0.0;
        for (int q = 0; q < comm.world; q++) s += all[(size_t)q * v.size() + i];
        v[i] = s;
    }
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: MIMSEM_RANK=r MIMSEM_WORLD=n host_dist_check sphere|box p ne nk rendezvous_dir\n");
        return 2;
    }
    const int kind = std::string(argv[1]) == "box" ? MIMSEM_MESH_BOX : MIMSEM_MESH_SPHERE;
    const int p = std::atoi(argv[2]), ne = std::atoi(argv[3]), nk = std::atoi(argv[4]);
    const int rank = std::atoi(getenv("MIMSEM_RANK") ? getenv("MIMSEM_RANK") : "0");
    const int world = std::atoi(getenv("MIMSEM_WORLD") ? getenv("MIMSEM_WORLD") : "1");
    const bool timing = argc > 6 && std::string(argv[6]) == "time";   // timing only: the parity checks run on the small meshes
    int failures = 0;
    try {
        GlobalMesh mesh;
        if (mesh.create(kind, p, ne)) throw std::runtime_error(mimsem_last_error());
        // layer thickness: a level profile times a horizontally non-uniform factor
        std::vector<double> thick((size_t)nk * mesh.NQ);
        for (int k = 0; k < nk; k++)
            for (int64_t q = 0; q < mesh.NQ; q++) {
                const double* x = &mesh.xyz[(size_t)q * 3];
                const double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
                thick[(size_t)k * mesh.NQ + q] = (200.0 + 30.0 * k) * (1.0 + 0.1 * (kind == MIMSEM_MESH_SPHERE ? x[2] / r : std::cos(x[0] * 6.283e-3)));
            }
        const std::vector<double> x1 = pseudo_random((size_t)nk * mesh.N1, 1, -1, 1), x2 = pseudo_random((size_t)nk * mesh.N2, 2, -1, 1);
        const std::vector<double> h2 = pseudo_random((size_t)nk * mesh.N2, 3, 0.5e4, 1.5e4), u1 = pseudo_random((size_t)nk * mesh.N1, 4, -1e9, 1e9);
        const std::vector<double> x0 = pseudo_random((size_t)nk * mesh.N0, 5, -1, 1), q0 = pseudo_random((size_t)nk * mesh.N0, 6, -1e-4, 1e-4);
        // advecting velocity of the upwinded operators: departure points stay well inside the element (tau = 150)
        double det_mean = 0.0;
        for (size_t i = 0; i < mesh.det.size(); i++) det_mean += std::fabs(mesh.det[i]) / (double)mesh.det.size();
        const std::vector<double> uu = pseudo_random((size_t)nk * mesh.N1, 7, -1e-3 * det_mean, 1e-3 * det_mean);
        FileComm comm(argv[5], rank, world);
        DistEngine eng(mesh, thick.data(), nk, &comm, rank);
        SelfComm self;
        DistEngine* one = (rank == 0 && !timing) ? new DistEngine(mesh, thick.data(), nk, &self, 0) : NULL;
        // all fifteen operators of the path (the set tests/mp_check.py runs through parallel.py)
        struct Case { const char* op; int sin, sout, sc; const std::vector<double>* x; const std::vector<double>* c; int tpow; bool up; };
        const Case cases[] = {{"M1", 1, 1, -1, &x1, NULL, 1, false},   {"M1h", 1, 1, 2, &x1, &h2, 2, false}, {"K", 1, 2, 1, &x1, &u1, 2, false},
                              {"M2", 2, 2, -1, &x2, NULL, 1, false},   {"M2h", 2, 2, 2, &x2, &h2, 2, false}, {"UtQW", 2, 1, 1, &x2, &u1, 0, false},
                              {"E21", 1, 2, -1, &x1, NULL, 0, false},  {"E12", 2, 1, -1, &x2, NULL, 0, false}, {"M0", 0, 0, -1, &x0, NULL, 1, false},
                              {"M0h", 0, 0, 2, &x0, &h2, 2, false},    {"E10", 0, 1, -1, &x0, NULL, 0, false}, {"E01", 1, 0, -1, &x1, NULL, 0, false},
                              {"R", 1, 1, 0, &x1, &q0, 2, false},      {"R_up", 1, 1, 0, &x1, &q0, 0, true},   {"M0h_up", 0, 0, 2, &x0, &h2, 0, true}};
        auto N_of = [&](int s) { return s == 0 ? mesh.N0 : (s == 1 ? mesh.N1 : mesh.N2); };
        auto run = [&](DistEngine& e, const Case& c, std::vector<double>& yg) {
            double* dx = e.alloc_field(c.sin, nk);
            double* dc = c.c ? e.alloc_field(c.sc, nk) : NULL;
            double* du = c.up ? e.alloc_field(1, nk) : NULL;
            double* dy = e.alloc_field(c.sout, nk);
            // owned rows from the global field, ghost rows ZERO: they must come from the exchange
            auto load = [&](const std::vector<double>& g, int space, double* d) {
                std::vector<double> own((size_t)nk * N_of(space), 0.0);
                const std::vector<int64_t>& ids = e.part().gids(space);
                for (int i = 0; i < e.part().n_owned(space); i++)
                    for (int k = 0; k < nk; k++) own[(size_t)k * N_of(space) + ids[i]] = g[(size_t)k * N_of(space) + ids[i]];
                e.scatter_from_global(own.data(), space, nk, d);
            };
            load(*c.x, c.sin, dx);
            if (dc) load(*c.c, c.sc, dc);
            if (du) load(uu, 1, du);
            for (int rep = 0; rep < 3; rep++) e.apply(c.op, dx, dc, dy, nk, 1.0e8, c.tpow, 0, du, c.up ? 150.0 : 0.0);
            e.sync();
            yg.assign((size_t)nk * N_of(c.sout), 0.0);
            e.owned_to_global(dy, c.sout, nk, yg.data());
            e.free_field(dx);
            if (dc) e.free_field(dc);
            if (du) e.free_field(du);
            e.free_field(dy);
        };
        for (size_t ci = 0; ci < sizeof(cases) / sizeof(cases[0]) && !timing; ci++) {
            std::vector<double> yg, ys;
            run(eng, cases[ci], yg);
            gather_sum(comm, yg);
            if (rank == 0) {
                run(*one, cases[ci], ys);
                const bool same = std::memcmp(yg.data(), ys.data(), yg.size() * 8) == 0;
                std::printf("%-6s on %d GPUs vs 1 GPU: %s\n", cases[ci].op, world, same ? "bitwise equal" : "DIFFERENT");
                if (!same) failures++;
            }
        }
        // partitioned solve: b = M1 x, then M1^-1 b recovers x; every rank stops at the same iteration
        if (!timing) {
            double* dx = eng.alloc_field(1, nk);
            double* db = eng.alloc_field(1, nk);
            double* ds = eng.alloc_field(1, nk);
            eng.scatter_from_global(x1.data(), 1, nk, dx);
            eng.apply_M1(dx, db, nk, 1.0e8, 1);
            double rr = 0.0;
            const int its = eng.solve_M1(db, ds, nk, 1.0e8, 1, 1e-13, 300, &rr);
            eng.sync();
            std::vector<double> xs((size_t)nk * mesh.N1, 0.0);
            eng.owned_to_global(ds, 1, nk, xs.data());
            gather_sum(comm, xs);
            double num = 0.0, den = 0.0;
            for (size_t i = 0; i < xs.size(); i++) {
                num += (xs[i] - x1[i]) * (xs[i] - x1[i]);
                den += x1[i] * x1[i];
            }
            std::vector<int> all_its(world);
            comm.allgather(&its, 4, all_its.data());
            bool agree = true;
            for (int q = 0; q < world; q++) agree = agree && all_its[q] == its;
            if (rank == 0) {
                std::printf("solve_M1 on %d GPUs: %d iterations, relres %.2e, error %.2e, ranks %s\n", world, its, rr, std::sqrt(num / den), agree ? "agree" : "DISAGREE");
                if (!(its < 300 && std::sqrt(num / den) < 1e-10 && agree)) failures++;
            }
            eng.free_field(dx);
            eng.free_field(db);
            eng.free_field(ds);
        }
        // bursts: six launches on two independent field pairs in one CUDA graph (programmatic dependencies between the
        // launches, consecutive epochs of the ghost hand-over), replayed three times; every pair bitwise equal to one GPU
        if (nk % 2 == 0) {
            const std::vector<double> x1b = pseudo_random((size_t)nk * mesh.N1, 9, -1, 1);
            const std::vector<double>* src[2] = {&x1, &x1b};
            std::vector<const double*> xs;
            std::vector<double*> ys;
            for (int i = 0; i < 2; i++) {
                double* dx = eng.alloc_field(1, nk);
                std::vector<double> own((size_t)nk * mesh.N1, 0.0);   // ghost rows zero: they must come over NVLink
                const std::vector<int64_t>& ids = eng.part().gids(1);
                for (int r = 0; r < eng.part().n_owned(1); r++)
                    for (int k = 0; k < nk; k++) own[(size_t)k * mesh.N1 + ids[r]] = (*src[i])[(size_t)k * mesh.N1 + ids[r]];
                eng.scatter_from_global(own.data(), 1, nk, dx);
                xs.push_back(dx);
                ys.push_back(eng.alloc_field(1, nk));
            }
            DistEngine::Burst* burst = eng.capture_burst_M1(xs, ys, 6, nk, 1.0e8, 1);
            for (int rep = 0; rep < 3; rep++) eng.replay(burst);
            eng.sync();
            for (int i = 0; i < 2 && !timing; i++) {
                std::vector<double> yg((size_t)nk * mesh.N1, 0.0);
                eng.owned_to_global(ys[i], 1, nk, yg.data());
                gather_sum(comm, yg);
                if (rank == 0) {
                    double* dx = one->alloc_field(1, nk);
                    double* dy = one->alloc_field(1, nk);
                    one->scatter_from_global(src[i]->data(), 1, nk, dx);
                    one->apply_M1(dx, dy, nk, 1.0e8, 1);
                    one->sync();
                    std::vector<double> y1((size_t)nk * mesh.N1, 0.0);
                    one->owned_to_global(dy, 1, nk, y1.data());
                    const bool same = std::memcmp(yg.data(), y1.data(), yg.size() * 8) == 0;
                    std::printf("burst of 6 x M1, field pair %d, on %d GPUs vs 1 GPU: %s\n", i, world, same ? "bitwise equal" : "DIFFERENT");
                    if (!same) failures++;
                    one->free_field(dx);
                    one->free_field(dy);
                }
            }
            if (timing) {
                // stream order (every launch waits for the one before) against bursts of 16 launches on three field pairs
                double* dx3 = eng.alloc_field(1, nk);
                eng.scatter_from_global(x1.data(), 1, nk, dx3);
                xs.push_back(dx3);
                ys.push_back(eng.alloc_field(1, nk));
                auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
                const int steps = 320;
                eng.sync(); comm.barrier();
                double t0 = now();
                for (int i = 0; i < steps; i++) eng.apply_M1(xs[i % 3], ys[i % 3], nk, 1.0e8, 1);
                eng.sync();
                double t_stream = now() - t0;
                DistEngine::Burst* b16 = eng.capture_burst_M1(xs, ys, 16, nk, 1.0e8, 1);
                for (int i = 0; i < 3; i++) eng.replay(b16);
                eng.sync(); comm.barrier();
                t0 = now();
                for (int i = 0; i < steps / 16; i++) eng.replay(b16);
                eng.sync();
                double t_burst = now() - t0;
                double tt[2] = {t_stream, t_burst};
                std::vector<double> all(2 * world);
                comm.allgather(tt, 16, all.data());
                for (int q = 0; q < world; q++) { t_stream = std::max(t_stream, all[2 * q]); t_burst = std::max(t_burst, all[2 * q + 1]); }
                if (rank == 0)
                    std::printf("{\"host\": \"C++ DistEngine, %d plain processes\", \"mesh\": \"%s p=%d ne=%d nk=%d\", \"steps\": %d, "
                                "\"stream_order_us_per_step\": %.2f, \"burst16_us_per_step\": %.2f, \"stream_order_gdofs\": %.1f, \"burst16_gdofs\": %.1f}\n",
                                world, argv[1], p, ne, nk, steps, t_stream / steps * 1e6, t_burst / steps * 1e6,
                                (double)mesh.N1 * nk * steps / t_stream / 1e9, (double)mesh.N1 * nk * steps / t_burst / 1e9);
                eng.free_burst(b16);
            }
            eng.free_burst(burst);
            for (size_t i = 0; i < xs.size(); i++) { eng.free_field(const_cast<double*>(xs[i])); eng.free_field(ys[i]); }
        }
        if (eng.halo_error()) {
            std::printf("rank %d: a ghost refresh timed out\n", rank);
            failures++;
        }
        comm.barrier();
        delete one;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rank %d: %s\n", rank, e.what());
        return 1;
    }
    if (rank == 0) std::printf("HOST_DIST_CHECK %s\n", failures ? "FAIL" : "OK");
    return failures ? 1 : 0;
}
