// Sanitizer driver for the kernel sources (TEST INFRASTRUCTURE; see tests/emu/cuda_emu.h).
//
// compute-sanitizer is closed on the GPU pool this project is measured on, so memcheck / racecheck /
// initcheck cannot be run on the device.  The substitute: the very same .cu / .cuh sources, compiled by g++
// against the CPU thread emulation (one OS thread per CUDA thread, pthread barriers for __syncthreads and for
// the warp collectives) with
//   -fsanitize=address,undefined   every shared-memory (heap block per CTA) and "device" buffer access is
//                                  bounds-checked - the memcheck analogue;
//   -fsanitize=thread              every pair of conflicting accesses by two emulated CUDA threads that is not
//                                  ordered by a barrier / shuffle is reported - the racecheck analogue (stricter:
//                                  it also covers global memory and code that would rely on warp lock-step);
// and shared memory is poisoned with 0xCD by the emulator, so a read of never-written shared memory shows up as
// NaN-like garbage in the checks below (initcheck analogue).  tools/run_cpu_sanitizers.sh builds and runs this.
//
// The workload mirrors tools/sanitize_sweep.py: two utterances, all four algorithms, n_fft 512 and 1024, three
// noise methods, an odd length, a NaN sample, clean caches, sweep, score expansion, selection, sparse enhance.
#include "../../classical_speech_enhancement_b200/csrc/cse_lib.cu"

#include <cassert>
#include <cstring>
#include <vector>

static std::vector<real> make_signal(int L, unsigned seed, const std::vector<real>* base, real noise) {
    std::vector<real> x(L);
    unsigned s = seed * 2654435761u + 12345u;
    for (int i = 0; i < L; ++i) {
        s = s * 1664525u + 1013904223u;
        const real r = (real)((s >> 8) & 0xffff) / R(65536) - R(0.5);
        const real tone = R(0.3) * (real)sin(2.0 * 3.14159265358979 * (180.0 + 40.0 * seed) * i / 16000.0) *
                          (real)(0.5 + 0.5 * sin(2.0 * 3.14159265358979 * 3.0 * i / 16000.0));
        x[i] = base ? (*base)[i] + noise * r : tone;
    }
    return x;
}

#define CHECK(call)                                                                     \
    do {                                                                                \
        const int rc_ = (call);                                                         \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, cse_last_error()); return 1; } \
    } while (0)

static int run_case(int L, int nan_at) {
    const int U = 2, sr = 16000;
    std::vector<unsigned char> tables(cse_tables_bytes());
    CHECK(cse_tables_init(tables.data(), nullptr));
    std::vector<real> clean, noisy;
    for (int u = 0; u < U; ++u) {
        std::vector<real> c = make_signal(L, 3 + u, nullptr, 0), n = make_signal(L, 17 + u, &c, R(0.2));
        clean.insert(clean.end(), c.begin(), c.end());
        noisy.insert(noisy.end(), n.begin(), n.end());
    }
    if (nan_at >= 0) noisy[(size_t)L + nan_at] = (real)NAN;
    std::vector<unsigned char> cache((size_t)U * cse_clean_cache_bytes(L, sr)), cws(cse_clean_workspace_bytes(U, L, sr));
    CHECK(cse_prepare_clean(tables.data(), clean.data(), U, L, sr, cache.data(), cws.data(), cws.size(), nullptr));
    long total = 0;
    const int shapes[2][2] = {{512, 128}, {1024, 256}};
    for (auto& sh : shapes) {
        const int n_fft = sh[0], hop = sh[1], nf = cse_num_frames(L, hop), nbp = cse_bins_padded(n_fft);
        std::vector<real> Y((size_t)U * nf * nbp * 2), P((size_t)U * nf * nbp), Nstat((size_t)U * nbp), Nmt((size_t)U * nf * nbp),
            Ntrue((size_t)U * nf * nbp);
        CHECK(cse_stft_psd(tables.data(), noisy.data(), nullptr, U, L, n_fft, hop, 0.0, Y.data(), P.data(), nullptr));
        std::vector<unsigned char> nws(cse_noise_workspace_bytes(U, nf, n_fft));
        CHECK(cse_noise_percentile(P.data(), U, nf, n_fft, 10.0, 1e-10, Nstat.data(), nws.data(), nws.size(), nullptr));
        CHECK(cse_noise_mintrack(P.data(), U, nf, n_fft, 1e-10, Nmt.data(), nullptr, 0, nullptr));
        CHECK(cse_stft_psd(tables.data(), noisy.data(), clean.data(), U, L, n_fft, hop, 1e-10, nullptr, Ntrue.data(), nullptr));
        const cse_params rows[4][2] = {
            {{{1.0, 0.01}}, {{4.0, 0.05}}},
            {{{0.95, 0.01}}, {{0.98, 0.1}}},
            {{{0.98, 0.001, 0.05, 1.0, 0.98}}, {{0.9, 0.1, 0.2, 1.0, -1.0}}},
            {{{0.9, 0.01, 0.1, 0.92, 0.3, 80.0}}, {{0.7, 0.001, 0.05, -1.0, 0.5, 80.0}}}};
        struct { const real* N; int tv; } noises[3] = {{Nstat.data(), 0}, {Nmt.data(), 1}, {Ntrue.data(), 1}};
        for (int alg = 0; alg < 4; ++alg)
            for (auto& nz : noises) {
                const int n_params = 2, chunk = 3;
                std::vector<cse_score_t> scores((size_t)U * n_params), nominal((size_t)U * 3);
                std::vector<unsigned char> ws(cse_sweep_workspace_bytes(chunk, L, sr));
                CHECK(cse_sweep(tables.data(), alg, Y.data(), nz.N, nz.tv, U, L, n_fft, hop, rows[alg], n_params, sr, clean.data(),
                                cache.data(), scores.data(), chunk, ws.data(), ws.size(), nullptr));
                const int base[3] = {0, 1, 0}, stride[3] = {2, 2, 2};       // three nominal points over two unique candidates
                CHECK(cse_expand_scores(scores.data(), base, stride, U, 3, nominal.data(), nullptr));
                std::vector<cse_winner_t> win((size_t)U * 3);
                const double pesq[6] = {2.0, 2.5, (double)NAN, 1.0, 3.0, 3.0004};
                CHECK(cse_select_best(nominal.data(), pesq, U, 3, win.data(), nullptr));
                if (alg >= 1) {      // the shared front: gamma in place of the PSD (noise_tv = 2) must give the same scores
                    std::vector<real> G((size_t)U * nf * nbp);
                    std::vector<cse_score_t> scores2((size_t)U * n_params);
                    const double mu = (alg == 2) ? rows[alg][0].v[4] : (alg == 3 ? rows[alg][0].v[3] : -1.0);
                    CHECK(cse_gamma(Y.data(), nz.N, nz.tv, U, L, n_fft, hop, nz.tv ? mu : -1.0, alg == 2 ? 1e-12 : 1e-10, G.data(), nullptr));
                    {   // grouped form: the same gamma twice in one launch == cse_gamma
                        std::vector<real> Ga((size_t)U * nf * nbp), Gb((size_t)U * nf * nbp);
                        const double e = alg == 2 ? 1e-12 : 1e-10, m = nz.tv ? mu : -1.0;
                        cse_gamma_group gg[2] = {{Y.data(), nz.N, Ga.data(), nz.tv, hop, m, e}, {Y.data(), nz.N, Gb.data(), nz.tv, hop, m, e}};
                        CHECK(cse_gamma_groups(U, L, n_fft, gg, 2, nullptr));
                        if (memcmp(G.data(), Ga.data(), G.size() * sizeof(real)) || memcmp(G.data(), Gb.data(), G.size() * sizeof(real))) {
                            fprintf(stderr, "grouped gamma differs: alg %d\n", alg);
                            return 1;
                        }
                    }
                    CHECK(cse_sweep(tables.data(), alg, Y.data(), G.data(), 2, U, L, n_fft, hop, rows[alg], 1, sr, clean.data(),
                                    cache.data(), scores2.data(), chunk, ws.data(), ws.size(), nullptr));
                    for (int u = 0; u < U; ++u)
                        if ((scores2[u].flags & CSE_FLAG_VALID) != (scores[(size_t)u * n_params].flags & CSE_FLAG_VALID) ||
                            ((scores2[u].flags & CSE_FLAG_VALID) && fabs((double)scores2[u].stoi - (double)scores[(size_t)u * n_params].stoi) > 1e-5)) {
                            fprintf(stderr, "gamma path differs: alg %d u %d: %g vs %g\n", alg, u, (double)scores2[u].stoi, (double)scores[(size_t)u * n_params].stoi);
                            return 1;
                        }
                }
                const int items[2] = {3, 0};
                std::vector<real> out((size_t)2 * L);
                CHECK(cse_enhance_list(tables.data(), alg, Y.data(), nz.N, nz.tv, L, n_fft, hop, rows[alg], n_params, items, 2,
                                       out.data(), nullptr));
                {   // grouped launch: two groups (the same inputs twice, second with one row) == cse_enhance of each
                    std::vector<real> ref((size_t)U * n_params * L), g0((size_t)U * n_params * L), g1((size_t)U * L);
                    CHECK(cse_enhance(tables.data(), alg, Y.data(), nz.N, nz.tv, U, L, n_fft, hop, rows[alg], n_params, ref.data(), nullptr));
                    cse_enhance_group gr[2] = {{Y.data(), nz.N, rows[alg], g0.data(), hop, n_params},
                                               {Y.data(), nz.N, rows[alg], g1.data(), hop, 1}};
                    CHECK(cse_enhance_groups(tables.data(), alg, nz.tv, U, L, n_fft, gr, 2, nullptr));
                    bool same = memcmp(ref.data(), g0.data(), ref.size() * sizeof(real)) == 0;
                    for (int u = 0; u < U && same; ++u)
                        same = memcmp(&ref[(size_t)u * n_params * L], &g1[(size_t)u * L], (size_t)L * sizeof(real)) == 0;
                    if (!same) { fprintf(stderr, "grouped launch differs: alg %d\n", alg); return 1; }
                }
                for (int u = 0; u < U; ++u)
                    for (int c = 0; c < n_params; ++c) {
                        const cse_score_t& s = scores[(size_t)u * n_params + c];
                        const bool expect_valid = !(nan_at >= 0 && u == 1);
                        if (((s.flags & CSE_FLAG_VALID) != 0) != expect_valid || (expect_valid && !(s.stoi > -1 && s.stoi <= 1.0001))) {
                            fprintf(stderr, "unexpected score alg %d u %d c %d: stoi %g flags %d\n", alg, u, c, (double)s.stoi, s.flags);
                            return 1;
                        }
                    }
                if (win[0].index < 0) { fprintf(stderr, "no winner\n"); return 1; }
                total += (long)U * n_params;
            }
    }
    printf("case L=%d nan_at=%d: %ld utterance-configs ok\n", L, nan_at, total);
    return 0;
}

// Negative controls: prove that the sanitizers see hazards in emulated kernels (run with "control-race" /
// "control-oob"; the sanitizer must then report).
__global__ void control_race_kernel(int* out) {
    CSE_DYN_SMEM(smem);
    int* a = reinterpret_cast<int*>(smem);
    a[threadIdx.x] = (int)threadIdx.x;
    /* missing __syncthreads() */
    out[threadIdx.x] = a[(threadIdx.x + 1) % blockDim.x];
}
__global__ void control_oob_kernel(int* out) {
    CSE_DYN_SMEM(smem);
    int* a = reinterpret_cast<int*>(smem);
    a[threadIdx.x] = 1;
    __syncthreads();
    out[threadIdx.x] = a[threadIdx.x + 2 * blockDim.x];      // past the 64 * 4 bytes the launch asked for
}

int main(int argc, char** argv) {
    if (argc > 1) {
        std::vector<int> out(64);
        if (!strcmp(argv[1], "control-race")) CSE_LAUNCH(control_race_kernel, 1, 64, 64 * sizeof(int), nullptr, out.data());
        if (!strcmp(argv[1], "control-oob")) CSE_LAUNCH(control_oob_kernel, 1, 64, 64 * sizeof(int), nullptr, out.data());
        printf("control %s done (out[0] = %d)\n", argv[1], out[0]);
        return 0;
    }
    if (run_case(4000, -1)) return 1;
    if (run_case(3501, 1700)) return 1;
    printf("sanitize_main ok\n");
    return 0;
}
