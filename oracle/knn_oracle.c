/* CPU oracle for the exhaustive inner-product search -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates faiss::IndexFlatIP::search as the reference uses it
 * (/root/reference/bioscanclip/util/util.py:522,525,528: index = IndexFlatIP(d);
 * index.add(keys); similarities, indices = index.search(queries, max_k)).
 * faiss (requirements.txt:22, faiss-gpu==1.7.2, CPU index) is a third-party
 * dependency absent from /root/reference: PARITY UNPINNED; see knn_oracle.py.
 *
 * Similarity of (q,k) = float64 sum over d = 0..D-1, in that order, of the exact
 * products (double)q[d]*(double)k[d] (a float32 x float32 product is exact in
 * float64).  Results ordered by (-sim, index): lowest index wins ties.
 * Compile WITHOUT -ffast-math so the summation order is kept.
 */
#include <stdint.h>
#include <stdlib.h>

int knn_oracle_search(const float* q, int64_t Q, const float* keys, int64_t K, int64_t D, int64_t k,
                      double* out_sims, int64_t* out_idx) {
    if (k <= 0 || k > K) return 1;
    int err = 0;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t qi = 0; qi < Q; ++qi) {
        const float* qv = q + qi * D;
        double* bs = out_sims + qi * k;
        int64_t* bi = out_idx + qi * k;
        int64_t have = 0;
        for (int64_t kj = 0; kj < K; ++kj) {
            const float* kv = keys + kj * D;
            double acc = 0.0;
            for (int64_t d = 0; d < D; ++d) acc += (double)qv[d] * (double)kv[d];
            /* keys arrive in increasing index, so a strict > keeps the lowest index on ties */
            if (have < k || acc > bs[have - 1]) {
                int64_t pos = have < k ? have : k - 1;
                while (pos > 0 && acc > bs[pos - 1]) {
                    bs[pos] = bs[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bs[pos] = acc;
                bi[pos] = kj;
                if (have < k) ++have;
            }
        }
    }
    return err;
}
