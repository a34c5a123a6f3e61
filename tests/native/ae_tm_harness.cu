// Stand-alone check + timing of the TMEM autoencoder kernels (st_ae_tm.cu) against a float64 CPU chain on random data.
// Test infrastructure (built by tests/native/Makefile, run under gpurun); not part of the product library.
//   ae_tm_harness [B=3] [C=8192] [K=4] [timing_B=200]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../signaltrain_b200/csrc/st_common.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static void make_dims(int C, int K, StDims& d, AeGeom& g) {
    const int N = 1024, H = 384;
    d.C = C; d.N = N; d.H = H; d.F = N / 2 + 1; d.K = K; d.R = 64;
    d.T = (C + H - 1) / H + (N + H - 1) / H;
    const int oc = C / 4;
    d.OT = (oc + H - 1) / H + (N + H - 1) / H;
    d.L = (d.OT - 1) * H - N;
    d.Fp = 528;
    d.Cp = C + 2 * N; d.Lp = d.L + 2 * N;
    d.Tp = (d.Cp + H - 1) / H; d.OTp = (d.Lp + H - 1) / H;
    d.Sx = d.Tp * H; d.Sg = d.OTp * H;
    const int in[ST_AE_LAYERS] = {d.T, 64, 32, 16, 16 + d.K, 16, 16, 32, 64};
    const int out[ST_AE_LAYERS] = {64, 32, 16, 16, 16, 16, 32, 64, d.OT};
    int flat = 0;
    for (int l = 0; l < ST_AE_LAYERS; ++l) {
        g.in[l] = in[l]; g.out[l] = out[l];
        g.flat_off[l] = flat;
        flat += out[l] * in[l] + out[l];
    }
    g.flat_total = flat;
}

struct HostAe {
    std::vector<float> W[ST_AE_LAYERS], b[ST_AE_LAYERS];
};

static double elu(double z) { return z > 0 ? z : std::expm1(z); }

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 3;
    const int C = argc > 2 ? atoi(argv[2]) : 8192;
    const int K = argc > 3 ? atoi(argv[3]) : 4;
    const int TB = argc > 4 ? atoi(argv[4]) : 200;
    StDims d; AeGeom g;
    make_dims(C, K, d, g);
    printf("geometry: C=%d T=%d OT=%d K=%d F=%d Fp=%d Tp=%d OTp=%d  B=%d (timing B=%d)\n", d.C, d.T, d.OT, d.K, d.F, d.Fp, d.Tp, d.OTp, B, TB);
    std::mt19937 rng(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::uniform_real_distribution<float> ud(-0.5f, 0.5f);
    HostAe ae[2];
    AeParams dp[2];
    for (int a = 0; a < 2; ++a)
        for (int l = 0; l < ST_AE_LAYERS; ++l) {
            const int n = g.in[l] * g.out[l];
            ae[a].W[l].resize(n); ae[a].b[l].resize(g.out[l]);
            const float sc = std::sqrt(2.f / (g.in[l] + g.out[l]));
            for (auto& w : ae[a].W[l]) w = sc * nd(rng);
            for (auto& w : ae[a].b[l]) w = 0.1f * nd(rng);
            float *dW, *db;
            CK(cudaMalloc(&dW, n * sizeof(float))); CK(cudaMalloc(&db, g.out[l] * sizeof(float)));
            CK(cudaMemcpy(dW, ae[a].W[l].data(), n * sizeof(float), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(db, ae[a].b[l].data(), g.out[l] * sizeof(float), cudaMemcpyHostToDevice));
            dp[a].W[l] = dW; dp[a].b[l] = db;
        }
    const int Bmax = std::max(B, TB);
    const long nspec = (long)Bmax * d.Tp * 2 * d.Fp, nri = (long)Bmax * d.OTp * 2 * d.Fp, nout = (long)Bmax * d.OT * d.F;
    std::vector<float> spec(nspec, 0.f), knobs((long)Bmax * std::max(K, 1));
    for (int b = 0; b < Bmax; ++b)
        for (int t = 0; t < d.T; ++t)
            for (int f = 0; f < d.F; ++f) {
                const long o = ((long)b * d.Tp + t) * 2 * d.Fp + f;
                // a few exactly-empty bins (zero-padding frames of the real model): mag = 0, phase = atan2(0, 1e-7) = 0
                const bool empty = (t == 0) || (f % 97 == 5);
                spec[o] = empty ? 0.f : 0.3f * nd(rng);
                spec[o + d.Fp] = empty ? 0.f : 0.3f * nd(rng);
            }
    for (auto& k : knobs) k = ud(rng);
    float *dspec, *dknobs, *dmag, *dmh, *dph, *dri, *dril, *ddbg, *dwpack;
    CK(cudaMalloc(&dwpack, st_ae_tm_pack_floats() * sizeof(float)));
    CK(cudaMalloc(&dspec, nspec * sizeof(float))); CK(cudaMalloc(&dknobs, knobs.size() * sizeof(float)));
    CK(cudaMalloc(&dmag, (long)Bmax * d.T * d.F * sizeof(float)));
    CK(cudaMalloc(&dmh, nout * sizeof(float))); CK(cudaMalloc(&dph, nout * sizeof(float)));
    CK(cudaMalloc(&dri, nri * sizeof(float))); CK(cudaMalloc(&dril, nri * sizeof(float)));
    CK(cudaMalloc(&ddbg, 2L * 9 * 128 * 64 * sizeof(float)));
    CK(cudaMemcpy(dspec, spec.data(), nspec * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dknobs, knobs.data(), knobs.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemset(dri, 0, nri * sizeof(float))); CK(cudaMemset(dril, 0, nri * sizeof(float)));
    CK(cudaMemset(ddbg, 0, 2L * 9 * 128 * 64 * sizeof(float)));
    int sm = 148;
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));

    // ---------------- forward check ----------------
    if (!st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, B, dmag, dmh, dph, dri, dril, dwpack, ddbg, nullptr, sm, true, 0, 0)) { printf("forward: geometry not covered\n"); return 3; }
    CK(cudaDeviceSynchronize());
    std::vector<float> mh(nout), ph(nout), ri(nri), ril(nri), mag((long)B * d.T * d.F), dbg(2L * 9 * 128 * 64);
    CK(cudaMemcpy(mh.data(), dmh, (long)B * d.OT * d.F * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ph.data(), dph, (long)B * d.OT * d.F * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ri.data(), dri, (long)B * d.OTp * 2 * d.Fp * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ril.data(), dril, (long)B * d.OTp * 2 * d.Fp * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(mag.data(), dmag, mag.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dbg.data(), ddbg, dbg.size() * sizeof(float), cudaMemcpyDeviceToHost));
    double e_mh = 0, e_ph = 0, e_ri = 0, e_mag = 0, e_lay[2][9] = {{0}};
    double m_mh = 0, m_ph = 0;
    for (int b = 0; b < B; ++b)
        for (int f = 0; f < d.F; ++f) {
            const long R = (long)b * d.F + f;
            double out[2][16];
            for (int a = 0; a < 2; ++a) {
                std::vector<double> h(64, 0.0), v(d.T);
                for (int t = 0; t < d.T; ++t) {
                    const long o = ((long)b * d.Tp + t) * 2 * d.Fp + f;
                    const double re = spec[o], im = spec[o + d.Fp];
                    v[t] = a == 0 ? std::sqrt(re * re + im * im) : std::atan2(im, (double)(float)(spec[o] + 1e-7f));
                    if (a == 0) e_mag = std::max(e_mag, std::fabs(v[t] - mag[((long)b * d.T + t) * d.F + f]));
                }
                std::vector<double> cur(v);
                for (int l = 0; l < ST_AE_LAYERS; ++l) {
                    if (l == 4) for (int k = 0; k < d.K; ++k) cur.push_back(knobs[(long)b * d.K + k]);
                    std::vector<double> nx(g.out[l]);
                    for (int o = 0; o < g.out[l]; ++o) {
                        double s = ae[a].b[l][o];
                        for (int i = 0; i < g.in[l]; ++i) s += (double)ae[a].W[l][o * g.in[l] + i] * cur[i];
                        nx[o] = elu(s);
                    }
                    if (R < 128)
                        for (int o = 0; o < g.out[l]; ++o)
                            e_lay[a][l] = std::max(e_lay[a][l], std::fabs(nx[o] - dbg[((long)(a * 9 + l) * 128 + R) * 64 + o]));
                    cur = nx;
                }
                for (int j = 0; j < d.OT; ++j) out[a][j] = a == 0 ? cur[j] * v[d.T - d.OT + j] : cur[j] + v[d.T - d.OT + j];
            }
            for (int j = 0; j < d.OT; ++j) {
                const long oo = ((long)b * d.OT + j) * d.F + f;
                e_mh = std::max(e_mh, std::fabs(out[0][j] - mh[oo])); m_mh = std::max(m_mh, std::fabs(out[0][j]));
                e_ph = std::max(e_ph, std::fabs(out[1][j] - ph[oo])); m_ph = std::max(m_ph, std::fabs(out[1][j]));
                const long orr = ((long)b * d.OTp + j) * 2 * d.Fp + f;
                e_ri = std::max(e_ri, std::fabs(out[0][j] * std::cos(out[1][j]) - ((double)ri[orr] + ril[orr])));
                e_ri = std::max(e_ri, std::fabs(out[0][j] * std::sin(out[1][j]) - ((double)ri[orr + d.Fp] + ril[orr + d.Fp])));
            }
        }
    printf("forward: max|mag err| %.3g  max|mag_hat err| %.3g (max %.3g)  max|phs_hat err| %.3g (max %.3g)  max|ri err| %.3g\n", e_mag, e_mh, m_mh, e_ph, m_ph, e_ri);
    for (int a = 0; a < 2; ++a) {
        printf("  layer errs ae%d:", a);
        for (int l = 0; l < 9; ++l) printf(" %.2g", e_lay[a][l]);
        printf("\n");
    }
    const bool fwd_ok = e_mh < 2e-5 && e_ph < 2e-5 && e_ri < 2e-5;
    printf("forward %s\n", fwd_ok ? "OK" : "MISMATCH");

    // ---------------- timing ----------------
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, TB, nullptr, dmh, dph, dri, dril, dwpack, nullptr, nullptr, sm, false, 0, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, TB, nullptr, dmh, dph, dri, dril, dwpack, nullptr, nullptr, sm, false, 0, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("forward timing: B=%d  %.2f us per launch (both autoencoders)\n", TB, 1000.f * ms / reps);
    {
        long long* dt;
        CK(cudaMalloc(&dt, 256 * sizeof(long long)));
        CK(cudaMemset(dt, 0, 256 * sizeof(long long)));
        st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, TB, nullptr, dmh, dph, dri, dril, dwpack, nullptr, dt, sm, false, 0, 0);
        CK(cudaDeviceSynchronize());
        long long ht[256];
        CK(cudaMemcpy(ht, dt, sizeof(ht), cudaMemcpyDeviceToHost));
        const long ntiles = ((long)TB * d.F + 127) / 128;
        const int grid = (int)std::min<long>(ntiles, sm);
        printf("forward regions (clk per CTA, %d CTAs, %.2f tiles per CTA):\n", grid, (double)ntiles / grid);
        printf("  issuer: wait %.0f  issue %.0f\n", (double)ht[4] / grid, (double)ht[5] / grid);
        for (int gidx = 0; gidx < 4; ++gidx)
            printf("  chain group %d (ae %d half %d): prologue %.0f  wait_d %.0f  epilogue %.0f  final %.0f\n", gidx, gidx / 2, gidx % 2,
                   (double)ht[8 * (1 + gidx)] / grid, (double)ht[8 * (1 + gidx) + 1] / grid, (double)ht[8 * (1 + gidx) + 2] / grid,
                   (double)ht[8 * (1 + gidx) + 3] / grid);
        // timeline of CTA 0, second tile: chain (arrive_l, wake_l) and issuer (seen_l, committed_l), relative clocks
        for (int a = 0; a < 2; ++a) {
            const long long t0 = ht[64 + a * 32];
            printf("  timeline ae%d:", a);
            for (int l = 0; l < 9; ++l)
                printf(" | L%d seen+%lld iss+%lld wake+%lld arr+%lld", l, ht[128 + a * 32 + 2 * l] - t0, ht[128 + a * 32 + 2 * l + 1] - t0,
                       ht[64 + a * 32 + 2 * l + 1] - t0, l < 8 ? ht[64 + a * 32 + 2 * l + 2] - t0 : 0LL);
            printf("\n");
        }
        printf("  epilogue probes ae0 (rel. to wake): ");
        for (int l = 0; l < 8; ++l)
            printf(" | L%d ld+%lld st_issued+%lld st_done+%lld arrive+%lld", l, ht[192 + 4 * l] - ht[64 + 2 * l + 1], ht[192 + 4 * l + 1] - ht[64 + 2 * l + 1],
                   ht[192 + 4 * l + 2] - ht[64 + 2 * l + 1], ht[64 + 2 * l + 2] - ht[64 + 2 * l + 1]);
        printf("\n");
    }
    return fwd_ok ? 0 : 1;
}
