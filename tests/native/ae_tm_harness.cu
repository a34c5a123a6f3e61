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
#define AE_DBG 0

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
    float* dtrk;                                             // [2][Bmax][T][F]: the forward kernel's tracks, read by the backward kernel
    CK(cudaMalloc(&dtrk, 2L * std::max(B, TB) * d.T * d.F * sizeof(float)));
    CK(cudaMemcpy(dspec, spec.data(), nspec * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dknobs, knobs.data(), knobs.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemset(dri, 0, nri * sizeof(float))); CK(cudaMemset(dril, 0, nri * sizeof(float)));
    CK(cudaMemset(ddbg, 0, 2L * 9 * 128 * 64 * sizeof(float)));
    int sm = 148;
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));

    // ---------------- forward check ----------------
    if (!st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, B, dmag, dtrk, dmh, dph, dri, dril, dwpack, ddbg, nullptr, sm, true, 0, 0)) { printf("forward: geometry not covered\n"); return 3; }
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

    // ---------------- backward check ----------------
    bool bwd_ok = true;
    if (d.T <= 32) {
        // CPU forward outputs as the kernel's inputs (isolates the backward), random upstream gradients
        std::vector<float> c_mh((long)B * d.OT * d.F), c_ph((long)B * d.OT * d.F), gri(nri, 0.f), gmh((long)B * d.OT * d.F);
        for (auto& v : gmh) v = 1e-3f * nd(rng);
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < d.OT; ++j)
                for (int f = 0; f < d.F; ++f) {
                    const long orr = ((long)b * d.OTp + j) * 2 * d.Fp + f;
                    gri[orr] = 1e-2f * nd(rng);
                    gri[orr + d.Fp] = 1e-2f * nd(rng);
                }
        std::vector<double> rW[2][9], rb[2][9];
        std::vector<double> rgt(2L * B * d.T * d.F, 0.0), rgz(18L * 128 * 64, 0.0);
        for (int a = 0; a < 2; ++a)
            for (int l = 0; l < 9; ++l) { rW[a][l].assign(g.in[l] * g.out[l], 0.0); rb[a][l].assign(g.out[l], 0.0); }
        const int tail0 = d.T - d.OT;
        // pass 1: forward of both autoencoders per row (mag_hat / phs_hat), pass 2: backward
        std::vector<std::vector<double>> acts(10);
        for (int pass = 0; pass < 2; ++pass)
            for (int b = 0; b < B; ++b)
                for (int f = 0; f < d.F; ++f) {
                    const long R = (long)b * d.F + f;
                    for (int a = 0; a < 2; ++a) {
                        acts[0].assign(d.T, 0.0);
                        for (int t = 0; t < d.T; ++t) {
                            const long o = ((long)b * d.Tp + t) * 2 * d.Fp + f;
                            const double re = spec[o], im = spec[o + d.Fp];
                            acts[0][t] = a == 0 ? std::sqrt(re * re + im * im) : std::atan2(im, (double)(float)(spec[o] + 1e-7f));
                        }
                        for (int l = 0; l < 9; ++l) {
                            std::vector<double> in(acts[l]);
                            if (l == 4) for (int k = 0; k < d.K; ++k) in.push_back(knobs[(long)b * d.K + k]);
                            if (l == 4) acts[4] = in;
                            acts[l + 1].assign(g.out[l], 0.0);
                            for (int o = 0; o < g.out[l]; ++o) {
                                double s2 = ae[a].b[l][o];
                                for (int i = 0; i < g.in[l]; ++i) s2 += (double)ae[a].W[l][o * g.in[l] + i] * in[i];
                                acts[l + 1][o] = elu(s2);
                            }
                        }
                        if (pass == 0) {
                            for (int j = 0; j < d.OT; ++j) {
                                const long oo = ((long)b * d.OT + j) * d.F + f;
                                if (a == 0) c_mh[oo] = (float)(acts[9][j] * acts[0][tail0 + j]);
                                else c_ph[oo] = (float)(acts[9][j] + acts[0][tail0 + j]);
                            }
                            continue;
                        }
                        std::vector<double> gz(d.OT), tb(d.OT);
                        for (int j = 0; j < d.OT; ++j) {
                            const long oo = ((long)b * d.OT + j) * d.F + f, orr = ((long)b * d.OTp + j) * 2 * d.Fp + f;
                            const double cs = std::cos((double)c_ph[oo]), sn = std::sin((double)c_ph[oo]);
                            const double e9 = acts[9][j], eg = e9 > 0 ? 1.0 : e9 + 1.0;
                            if (a == 0) {
                                const double gm = gri[orr] * cs + gri[orr + d.Fp] * sn + gmh[oo];
                                gz[j] = gm * acts[0][tail0 + j] * eg; tb[j] = gm * e9;
                            } else {
                                const double gp = (double)c_mh[oo] * (gri[orr + d.Fp] * cs - gri[orr] * sn);
                                gz[j] = gp * eg; tb[j] = gp;
                            }
                        }
                        for (int l = 8; l >= 0; --l) {
                            if (a == AE_DBG && R < 128) for (int o = 0; o < g.out[l]; ++o) rgz[((long)(9 + l) * 128 + R) * 64 + o] = gz[o];
                            const std::vector<double>& in = acts[l];
                            for (int o = 0; o < g.out[l]; ++o) {
                                rb[a][l][o] += gz[o];
                                for (int i = 0; i < g.in[l]; ++i) rW[a][l][o * g.in[l] + i] += gz[o] * in[i];
                            }
                            const int nin = l == 4 ? 16 : g.in[l];
                            std::vector<double> gh(nin, 0.0);
                            for (int i = 0; i < nin; ++i)
                                for (int o = 0; o < g.out[l]; ++o) gh[i] += gz[o] * ae[a].W[l][o * g.in[l] + i];
                            if (l > 0) {
                                gz.assign(nin, 0.0);
                                for (int i = 0; i < nin; ++i) gz[i] = gh[i] * (acts[l][i] > 0 ? 1.0 : acts[l][i] + 1.0);
                            } else {
                                for (int t = 0; t < d.T; ++t)
                                    rgt[(long)a * B * d.T * d.F + ((long)b * d.T + t) * d.F + f] = gh[t] + (t >= tail0 ? tb[t - tail0] : 0.0);
                            }
                        }
                    }
                }
        float *dcmh, *dcph, *dgri, *dgmh, *dgt, *dpart, *dwb, *ddbg2;
        const long ntrk = (long)B * d.T * d.F;
        CK(cudaMalloc(&dcmh, c_mh.size() * 4)); CK(cudaMalloc(&dcph, c_ph.size() * 4)); CK(cudaMalloc(&dgri, std::max<long>(nri, 1) * 4));
        CK(cudaMalloc(&dgmh, (long)Bmax * d.OT * d.F * 4)); CK(cudaMalloc(&dgt, 2L * Bmax * d.T * d.F * 4));
        CK(cudaMalloc(&dpart, (long)sm * g.flat_total * 4)); CK(cudaMalloc(&dwb, st_ae_tm_bwd_pack_floats() * 4)); CK(cudaMalloc(&ddbg2, 18L * 128 * 64 * 4));
        CK(cudaMemcpy(dcmh, c_mh.data(), c_mh.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcph, c_ph.data(), c_ph.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dgri, gri.data(), nri * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dgmh, 0, (long)Bmax * d.OT * d.F * 4));
        CK(cudaMemcpy(dgmh, gmh.data(), gmh.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dpart, 0, (long)sm * g.flat_total * 4)); CK(cudaMemset(ddbg2, 0, 18L * 128 * 64 * 4)); CK(cudaMemset(dgt, 0, 2L * Bmax * d.T * d.F * 4));
        const int nslot = st_launch_ae_backward_tm(d, g, dp[0], dp[1], dspec, dtrk, dknobs, B, dcmh, dcph, dgri, dgmh, nullptr, dgt, nullptr, nullptr, dpart, dwb,
                                                   ddbg2, nullptr, sm, true, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("backward: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        printf("backward: %d partial slots per autoencoder\n", nslot);
        std::vector<float> gt(2 * ntrk), part((long)2 * nslot * g.flat_total), dbg2(18L * 128 * 64);
        CK(cudaMemcpy(gt.data(), dgt, gt.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(part.data(), dpart, part.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(dbg2.data(), ddbg2, dbg2.size() * 4, cudaMemcpyDeviceToHost));
        for (int a = 0; a < 2; ++a) {
            double eg = 0, mg = 0;
            for (long i = 0; i < ntrk; ++i) { eg = std::max(eg, std::fabs(rgt[a * ntrk + i] - gt[a * ntrk + i])); mg = std::max(mg, std::fabs(rgt[a * ntrk + i])); }
            printf("  ae%d track gradient: max err %.3g of max %.3g (%.2g rel)\n", a, eg, mg, eg / mg);
            if (eg > 3e-4 * mg) bwd_ok = false;
            for (int l = 0; l < 9; ++l) {
                double ew = 0, mw = 0, eb = 0, mb = 0;
                for (int i = 0; i < g.in[l] * g.out[l]; ++i) {
                    double sum = 0;
                    for (int sidx = 0; sidx < nslot; ++sidx) sum += part[((long)sidx * 2 + a) * g.flat_total + g.flat_off[l] + i];
                    ew = std::max(ew, std::fabs(sum - rW[a][l][i])); mw = std::max(mw, std::fabs(rW[a][l][i]));
                }
                for (int i = 0; i < g.out[l]; ++i) {
                    double sum = 0;
                    for (int sidx = 0; sidx < nslot; ++sidx) sum += part[((long)sidx * 2 + a) * g.flat_total + g.flat_off[l] + g.in[l] * g.out[l] + i];
                    eb = std::max(eb, std::fabs(sum - rb[a][l][i])); mb = std::max(mb, std::fabs(rb[a][l][i]));
                }
                printf("    layer %d: dW err %.3g / %.3g (%.2g)   db err %.3g / %.3g (%.2g)\n", l, ew, mw, ew / mw, eb, mb, eb / mb);
                if (ew > 3e-4 * mw || eb > 3e-4 * mb) bwd_ok = false;
            }
        }
        {
            printf("  tile-0 chain of ae%d (dbg): fwd layer errs:", AE_DBG);
            // the dbg buffer holds the chain of whichever autoencoder's CTA 0/1 processed tile 0: both write the same slots, the
            // kernel lets autoencoder AE_DBG win by launching order being unspecified -> compare against the closer one
            for (int l = 0; l < 9; ++l) {
                double em = 0;
                for (int r = 0; r < 128 && r < B * d.F; ++r) for (int o = 0; o < g.out[l]; ++o) em = std::max(em, std::fabs(rgz[((long)(9 + l) * 128 + r) * 64 + o] - dbg2[((long)(9 + l) * 128 + r) * 64 + o]));
                printf(" gz%d %.2g", l, em);
            }
            printf("\n");
        }
        printf("backward %s\n", bwd_ok ? "OK" : "MISMATCH");
        // timing
        cudaEvent_t b0, b1;
        CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&b1));
        float *dcmh2, *dcph2;
        CK(cudaMalloc(&dcmh2, nout * 4)); CK(cudaMalloc(&dcph2, nout * 4));
        CK(cudaMemset(dcmh2, 0, nout * 4)); CK(cudaMemset(dcph2, 0, nout * 4));
        float* dgri2; CK(cudaMalloc(&dgri2, nri * 4)); CK(cudaMemset(dgri2, 0, nri * 4));
        st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, TB, nullptr, dtrk, dmh, dph, dri, dril, dwpack, nullptr, nullptr, sm, false, 0, 0);
        for (int i = 0; i < 3; ++i) st_launch_ae_backward_tm(d, g, dp[0], dp[1], dspec, dtrk, dknobs, TB, dcmh2, dcph2, dgri2, dgmh, nullptr, dgt, nullptr, nullptr, dpart, dwb, nullptr, nullptr, sm, false, 0, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(b0));
        for (int i = 0; i < 20; ++i) st_launch_ae_backward_tm(d, g, dp[0], dp[1], dspec, dtrk, dknobs, TB, dcmh2, dcph2, dgri2, dgmh, nullptr, dgt, nullptr, nullptr, dpart, dwb, nullptr, nullptr, sm, false, 0, 0);
        CK(cudaEventRecord(b1));
        CK(cudaDeviceSynchronize());
        float bms = 0;
        CK(cudaEventElapsedTime(&bms, b0, b1));
        printf("backward timing: B=%d  %.2f us per launch (both autoencoders, without the track->spec kernel)\n", TB, 1000.f * bms / 20);
        {
            long long* dt2;
            CK(cudaMalloc(&dt2, 128 * sizeof(long long)));
            CK(cudaMemset(dt2, 0, 128 * sizeof(long long)));
            const int ns = st_launch_ae_backward_tm(d, g, dp[0], dp[1], dspec, dtrk, dknobs, TB, dcmh2, dcph2, dgri2, dgmh, nullptr, dgt, nullptr, nullptr, dpart, dwb, nullptr, dt2, sm, false, 0, 0);
            CK(cudaDeviceSynchronize());
            long long ht2[128];
            CK(cudaMemcpy(ht2, dt2, sizeof(ht2), cudaMemcpyDeviceToHost));
            const char* names[7] = {"prologue", "wait_d", "fwd_epi", "dec", "bwd_epi", "handoff+stage", "final"};
            printf("backward chain-warp clocks per CTA (%d CTAs):\n", 2 * ns);
            for (int w = 0; w < 8; w += 3) {
                printf("  warp %d:", w);
                for (int i = 0; i < 7; ++i) printf(" %s %.0f", names[i], (double)ht2[8 * w + i] / (2 * ns));
                printf("\n           handoff: tmem %.0f wait_free %.0f gz_sts %.0f act %.0f fence+full %.0f\n", (double)ht2[64 + 8 * w] / (2 * ns),
                       (double)ht2[64 + 8 * w + 1] / (2 * ns), (double)ht2[64 + 8 * w + 2] / (2 * ns), (double)ht2[64 + 8 * w + 3] / (2 * ns),
                       (double)ht2[64 + 8 * w + 4] / (2 * ns));
            }
        }
    }

    // ---------------- timing ----------------
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, TB, nullptr, dtrk, dmh, dph, dri, dril, dwpack, nullptr, nullptr, sm, false, 0, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, TB, nullptr, dtrk, dmh, dph, dri, dril, dwpack, nullptr, nullptr, sm, false, 0, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("forward timing: B=%d  %.2f us per launch (both autoencoders)\n", TB, 1000.f * ms / reps);
    {
        long long* dt;
        CK(cudaMalloc(&dt, 256 * sizeof(long long)));
        CK(cudaMemset(dt, 0, 256 * sizeof(long long)));
        st_launch_ae_forward_tm(d, g, dp[0], dp[1], dspec, dknobs, TB, nullptr, dtrk, dmh, dph, dri, dril, dwpack, nullptr, dt, sm, false, 0, 0);
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
    return (fwd_ok && bwd_ok) ? 0 : 1;
}
