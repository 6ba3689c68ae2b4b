/*
 * oracle/knn_recipe.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Scalar C restatement of the arithmetic the reference's kNN graph build performs on
 * the CPU, i.e. Generation/modules.py:695-704 of liruihui/SP-GAN:
 *     xi   = -2 * bmm(x^T, x)                      (modules.py:696)
 *     xs   = sum(x^T ** 2, dim=2, keepdim=True)    (modules.py:697)
 *     dist = xi + xs + xs^T                        (modules.py:699)
 *     idx  = sort(dist, dim=2)[..., 1:k+1]         (modules.py:702-703)
 * The arithmetic itself lives in PyTorch (not vendored by the reference).  The
 * rounding order restated here is the one torch 2.11 CPU (MKL sgemm + ATen cascade
 * sum) executes; tests/test_oracle_knn.py pins it bit-for-bit against torch on the
 * build host and against the golden vectors produced by the unmodified reference.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -march=x86-64-v3 (see oracle/Makefile).
 * -ffp-contract=off matters: only the explicit fmaf() calls may fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int ceil_log2_i64(int64_t x) {
    if (x <= 2) return 1;
    int r = 0;
    uint64_t v = (uint64_t)(x - 1);
    while (v) { v >>= 1; ++r; }
    return r;
}

/* Sum of squares over the channel axis of one point: products rounded to fp32
 * first (the reference materialises x**2), then the multi-level cascade that ATen's
 * strided-reduction kernel uses: chunks of 2^p rows are summed sequentially into
 * level 0, every full chunk is folded into the next level, levels are finally added
 * 0+1, +2, +3. */
static float sqnorm_cascade(const float *x, int64_t C, int64_t stride) {
    enum { LEVELS = 4 };
    int p = ceil_log2_i64(C) / LEVELS;
    if (p < 4) p = 4;
    const int64_t step = (int64_t)1 << p;
    const int64_t mask = step - 1;
    float acc[LEVELS] = {0.f, 0.f, 0.f, 0.f};
    int64_t i = 0;
    while (i + step <= C) {
        for (int64_t j = 0; j < step; ++j, ++i) {
            const float v = x[i * stride];
            const float sq = v * v;      /* rounded product, no FMA */
            acc[0] = acc[0] + sq;
        }
        for (int l = 1; l < LEVELS; ++l) {
            acc[l] = acc[l] + acc[l - 1];
            acc[l - 1] = 0.f;
            const int64_t m = mask << (l * p);
            if ((i & m) != 0) break;
        }
    }
    for (; i < C; ++i) {
        const float v = x[i * stride];
        const float sq = v * v;
        acc[0] = acc[0] + sq;
    }
    for (int l = 1; l < LEVELS; ++l) acc[0] = acc[0] + acc[l];
    return acc[0];
}

/* The same reduction for a column that ATen handles outside its 4-vector-wide main
 * loop: the channel axis is viewed as [C/4, 4], four interleaved partial cascades are
 * kept (partial q sums channels 4i+q), leftover channels go to partial 0, and the
 * partials are folded 0+1, +2, +3. */
static float sqnorm_ilp4(const float *x, int64_t C, int64_t stride) {
    enum { LEVELS = 4, ILP = 4 };
    const int64_t G = C / ILP;
    int p = ceil_log2_i64(G) / LEVELS;
    if (p < 4) p = 4;
    const int64_t step = (int64_t)1 << p;
    const int64_t mask = step - 1;
    float acc[LEVELS][ILP];
    memset(acc, 0, sizeof(acc));
    int64_t i = 0;
    while (i + step <= G) {
        for (int64_t j = 0; j < step; ++j, ++i)
            for (int q = 0; q < ILP; ++q) {
                const float v = x[(i * ILP + q) * stride];
                const float sq = v * v;
                acc[0][q] = acc[0][q] + sq;
            }
        for (int l = 1; l < LEVELS; ++l) {
            for (int q = 0; q < ILP; ++q) {
                acc[l][q] = acc[l][q] + acc[l - 1][q];
                acc[l - 1][q] = 0.f;
            }
            const int64_t m = mask << (l * p);
            if ((i & m) != 0) break;
        }
    }
    for (; i < G; ++i)
        for (int q = 0; q < ILP; ++q) {
            const float v = x[(i * ILP + q) * stride];
            const float sq = v * v;
            acc[0][q] = acc[0][q] + sq;
        }
    for (int l = 1; l < LEVELS; ++l)
        for (int q = 0; q < ILP; ++q) acc[0][q] = acc[0][q] + acc[l][q];
    for (int64_t c = G * ILP; c < C; ++c) {
        const float v = x[c * stride];
        const float sq = v * v;
        acc[0][0] = acc[0][0] + sq;
    }
    for (int q = 1; q < ILP; ++q) acc[0][0] = acc[0][0] + acc[0][q];
    return acc[0][0];
}

/* x: [B, C, N] channel-first fp32 (the reference layout).  xs: [B, N].
 * main_cols: number of leading columns n < main_cols that use the plain cascade;
 * the rest use the ILP-4 variant.  The reference arithmetic on an AVX-512 host is
 * main_cols = (N / 32) * 32 (ATen outer-sum main loop: 4 AVX2 vectors of 8 floats);
 * pass a negative value for that default.  Every N the reference ships a sphere
 * template for (256 ... 20000) is a multiple of 32, i.e. pure cascade. */
void spgan_oracle_sqnorm_ex(const float *x, int B, int C, int N, int main_cols, float *xs) {
    if (main_cols < 0) main_cols = (N / 32) * 32;
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n) {
            const float *col = x + (int64_t)b * C * N + n;
            xs[(int64_t)b * N + n] = (n < main_cols) ? sqnorm_cascade(col, C, N)
                                                     : sqnorm_ilp4(col, C, N);
        }
}

void spgan_oracle_sqnorm(const float *x, int B, int C, int N, float *xs) {
    spgan_oracle_sqnorm_ex(x, B, C, N, -1, xs);
}

/* One entry of the reference `dist` matrix, in the reference's rounding order. */
static inline float dist_entry(const float *xb, int C, int N, int i, int j,
                               float xs_i, float xs_j) {
    float dot = 0.f;
    for (int c = 0; c < C; ++c)
        dot = fmaf(xb[(int64_t)c * N + i], xb[(int64_t)c * N + j], dot);
    const float xi = -2.0f * dot;        /* exact scaling */
    const float t = xi + xs_i;           /* (xi + xs) ...  */
    return t + xs_j;                     /* ... + xs^T     */
}

/* Full [B, N, N] distance matrix (small shapes only). */
void spgan_oracle_dist(const float *x, int B, int C, int N, float *dist) {
    float *xs = (float *)malloc(sizeof(float) * (size_t)B * N);
    spgan_oracle_sqnorm(x, B, C, N, xs);
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (int64_t)b * C * N;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                dist[((int64_t)b * N + i) * N + j] =
                    dist_entry(xb, C, N, i, j, xs[(int64_t)b * N + i], xs[(int64_t)b * N + j]);
    }
    free(xs);
}

/* Ranks 1..k of every row in ascending (dist, j) order (rank 0 is dropped whatever
 * it is -- modules.py:703).  idx: [B, N, k] int32.  kdist (optional, may be NULL):
 * [B, N, k+1] the k+1 smallest distances including rank 0, for tie analysis.
 * Returns 0, or -1 on bad arguments. */
int spgan_oracle_knn(const float *x, int B, int C, int N, int k, int32_t *idx, float *kdist) {
    if (B < 0 || C < 1 || N < 1 || k < 1 || k + 1 > N) return -1;
    const int K1 = k + 1;
    float *xs = (float *)malloc(sizeof(float) * (size_t)B * N);
    float *bd = (float *)malloc(sizeof(float) * (size_t)K1);
    int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)K1);
    float *xi_col = (float *)malloc(sizeof(float) * (size_t)C);
    spgan_oracle_sqnorm(x, B, C, N, xs);
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (int64_t)b * C * N;
        const float *xsb = xs + (int64_t)b * N;
        for (int i = 0; i < N; ++i) {
            int cnt = 0;
            for (int c = 0; c < C; ++c) xi_col[c] = xb[(int64_t)c * N + i];
            for (int j = 0; j < N; ++j) {
                float dot = 0.f;
                for (int c = 0; c < C; ++c) dot = fmaf(xi_col[c], xb[(int64_t)c * N + j], dot);
                const float d = (-2.0f * dot + xsb[i]) + xsb[j];
                /* candidates arrive in increasing j: strict '<' keeps (dist, j) order */
                if (cnt == K1 && !(d < bd[K1 - 1])) continue;
                int pos = (cnt < K1) ? cnt : K1 - 1;
                while (pos > 0 && d < bd[pos - 1]) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = j;
                if (cnt < K1) ++cnt;
            }
            for (int r = 0; r < k; ++r) idx[((int64_t)b * N + i) * k + r] = bi[r + 1];
            if (kdist)
                for (int r = 0; r < K1; ++r) kdist[((int64_t)b * N + i) * K1 + r] = bd[r];
        }
    }
    free(xs); free(bd); free(bi); free(xi_col);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * The reference's other distance entry points: two clouds, either norm order.
 *   knn            Generation/modules.py:640-646   pd = -xx - inner - xx^T   (candidate norm first, negated)
 *   pairwise_dist  Generation/modules.py:629-637   dist = xy + xx + yy^T     (query norm first)
 *   square_distance Common/pointnet_util.py:19-40  dist = -2 src.dst^T; += |src|^2; += |dst|^2
 * xq: [B, C, Nq], xc: [B, C, Nc] channel-first; xsq [B, Nq], xsc [B, Nc] squared norms supplied by the caller
 * (the reduction order of the norms depends on the memory layout the reference reduces over).
 * dist: [B, Nq, Nc].  The dot product is the same FMA chain over channels as above. */
void spgan_oracle_dist2(const float *xq, const float *xsq, int Nq, const float *xc, const float *xsc, int Nc,
                        int B, int C, int cand_norm_first, float *dist) {
    for (int b = 0; b < B; ++b) {
        const float *qb = xq + (int64_t)b * C * Nq;
        const float *cb = xc + (int64_t)b * C * Nc;
        for (int i = 0; i < Nq; ++i)
            for (int j = 0; j < Nc; ++j) {
                float dot = 0.f;
                for (int c = 0; c < C; ++c) dot = fmaf(qb[(int64_t)c * Nq + i], cb[(int64_t)c * Nc + j], dot);
                const float m2 = -2.0f * dot;
                const float nq = xsq[(int64_t)b * Nq + i], nc = xsc[(int64_t)b * Nc + j];
                float d;
                if (cand_norm_first) { const float t = m2 + nc; d = t + nq; }
                else { const float t = m2 + nq; d = t + nc; }
                dist[((int64_t)b * Nq + i) * Nc + j] = d;
            }
    }
}

/* Point-major squared norms, rounded squares added in channel order (what torch CPU computes for
 * `torch.sum(src ** 2, -1)` on xyz rows, C <= 3). */
void spgan_oracle_sqnorm_rows(const float *x, int64_t R, int C, float *xs) {
    for (int64_t r = 0; r < R; ++r) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = x[r * C + c];
            const float sq = v * v;
            acc = acc + sq;
        }
        xs[r] = acc;
    }
}

/* Ranks first_rank .. first_rank+k-1 of every row of a [R, Nc] distance matrix in ascending (dist, j) order.
 * idx: [R, k] int32.  Returns 0, or -1 on bad arguments. */
int spgan_oracle_topk_rows(const float *dist, int64_t R, int Nc, int k, int first_rank, int32_t *idx) {
    if (R < 0 || Nc < 1 || k < 1 || first_rank < 0 || k + first_rank > Nc) return -1;
    const int K1 = k + first_rank;
    float *bd = (float *)malloc(sizeof(float) * (size_t)K1);
    int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)K1);
    for (int64_t r = 0; r < R; ++r) {
        int cnt = 0;
        for (int j = 0; j < Nc; ++j) {
            const float d = dist[r * Nc + j];
            if (cnt == K1 && !(d < bd[K1 - 1])) continue;
            int pos = (cnt < K1) ? cnt : K1 - 1;
            while (pos > 0 && d < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
            bd[pos] = d;
            bi[pos] = j;
            if (cnt < K1) ++cnt;
        }
        for (int t = 0; t < k; ++t) idx[r * k + t] = bi[t + first_rank];
    }
    free(bd); free(bi);
    return 0;
}
