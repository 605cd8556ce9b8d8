// Stand-alone check of kc::merge_pair (kc_merge.cu) against std::merge on random sorted unique runs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I../../kmer-counter_b200/csrc merge_test.cu \
//        ../../kmer-counter_b200/csrc/kc_merge.cu -o merge_test
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>

#include "kc_internal.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

typedef std::vector<uint64_t> KeyV;

#include <chrono>
#include <unistd.h>
int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    std::mt19937_64 rng(7);
    for (int W = 1; W <= 4; W++) {
        for (int trial = 0; trial < 12; trial++) {
            const size_t na = trial == 0 ? 0 : (rng() % 40000), nb = trial == 1 ? 0 : (rng() % 40000) + 1;
            const uint64_t range = trial % 3 == 0 ? 5000 : 1000000;          // small range: many equal pairs
            auto gen = [&](size_t n) {
                std::map<KeyV, uint32_t> m;
                while (m.size() < n && m.size() < range) {
                    KeyV k(W);
                    for (int i = 0; i < W; i++) k[i] = i + 1 == W ? rng() % range : (rng() % 2);
                    m[k] = (uint32_t)(rng() % 1000) + 1;
                }
                return m;
            };
            auto A = gen(na), B = gen(nb);
            std::map<KeyV, uint32_t> want = A;
            for (auto &kv : B) want[kv.first] += kv.second;
            auto flat = [&](const std::map<KeyV, uint32_t> &m, std::vector<uint64_t> &k, std::vector<uint32_t> &c) {
                for (auto &kv : m) { for (int i = 0; i < W; i++) k.push_back(kv.first[i]); c.push_back(kv.second); }
            };
            std::vector<uint64_t> ka, kb, kw;
            std::vector<uint32_t> ca, cb, cw;
            flat(A, ka, ca); flat(B, kb, cb); flat(want, kw, cw);
            const size_t nA = ca.size(), nB = cb.size();
            uint64_t *dka, *dkb, *dko;
            uint32_t *dca, *dcb, *dco;
            unsigned long long *dn;
            void *ws;
            CK(cudaMalloc(&dka, nA * W * 8 + 16)); CK(cudaMalloc(&dkb, nB * W * 8 + 16)); CK(cudaMalloc(&dko, (nA + nB) * W * 8 + 16));
            CK(cudaMalloc(&dca, nA * 4 + 16)); CK(cudaMalloc(&dcb, nB * 4 + 16)); CK(cudaMalloc(&dco, (nA + nB) * 4 + 16));
            CK(cudaMalloc(&dn, 8)); CK(cudaMalloc(&ws, kc::merge_workspace_bytes(nA, nB)));
            CK(cudaMemcpy(dka, ka.data(), nA * W * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dkb, kb.data(), nB * W * 8, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dca, ca.data(), nA * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcb, cb.data(), nB * 4, cudaMemcpyHostToDevice));
            int launches = 0;
            printf("W=%d na=%zu nb=%zu ... ", W, nA, nB);
            cudaStream_t st;
            CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            CK(kc::merge_pair(dka, dca, nA, dkb, dcb, nB, W, dko, dco, dn, ws, st, &launches));
            auto t0 = std::chrono::steady_clock::now();
            while (cudaStreamQuery(st) == cudaErrorNotReady) {
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 5.0) {
                    unsigned long long *hws = nullptr;
                    printf("HANG\n");
                    _exit(3);
                }
                usleep(1000);
            }
            CK(cudaStreamSynchronize(st));
            unsigned long long U = 0;
            CK(cudaMemcpy(&U, dn, 8, cudaMemcpyDeviceToHost));
            std::vector<uint64_t> ko(U * W);
            std::vector<uint32_t> co(U);
            CK(cudaMemcpy(ko.data(), dko, U * W * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(co.data(), dco, U * 4, cudaMemcpyDeviceToHost));
            const bool ok = U == cw.size() && ko == kw && co == cw;
            printf("W=%d na=%zu nb=%zu U=%llu want=%zu %s\n", W, nA, nB, U, cw.size(), ok ? "ok" : "MISMATCH");
            if (!ok) {
                int shown = 0;
                for (size_t i = 0; i < U && i < cw.size() && shown < 12; i++) {
                    bool same = co[i] == cw[i];
                    for (int q = 0; q < W; q++) same = same && ko[i * W + q] == kw[i * W + q];
                    if (!same) {
                        printf("  record %zu: got", i);
                        for (int q = 0; q < W; q++) printf(" %llx", (unsigned long long)ko[i * W + q]);
                        printf(" :%u   want", co[i]);
                        for (int q = 0; q < W; q++) printf(" %llx", (unsigned long long)kw[i * W + q]);
                        printf(" :%u\n", cw[i]);
                        shown++;
                    }
                }
                return 2;
            }
            cudaFree(dka); cudaFree(dkb); cudaFree(dko); cudaFree(dca); cudaFree(dcb); cudaFree(dco); cudaFree(dn); cudaFree(ws);
        }
    }
    printf("all ok\n");
    return 0;
}
