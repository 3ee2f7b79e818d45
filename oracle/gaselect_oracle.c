/* oracle/gaselect_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C CPU restatement of the reference's (dakep/gaselect) fitness hot path, written
 * from the algorithm, citing the reference file:line each function follows.  It is the
 * checker the CUDA engine is compared with on the GPU box (where /root/reference does not
 * exist) and the "port" CPU baseline of bench.py.
 *
 * PARITY PIN: this restatement is pinned against (a) the known-answer vectors of
 * SURVEY.md Appendix C, committed under tests/golden/, and (b) live cross-execution with
 * oracle/_ref/libgaselect_ref.so -- the reference's own C++ compiled out of tree -- in
 * tests/test_oracle.py.  The reference itself ships no tests or golden vectors.
 * Third-party arithmetic not in the reference tree: arma::solve (LAPACK least squares) is
 * restated here as Householder QR; agreement is to rounding (1e-13) for full-rank systems.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 * The entry points mirror oracle/ref_shim.cpp (prefix orc_ instead of ref_).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * WELL19937a with the reference's tempering and index handling
 * follows src/RNG.cpp:61-167 and src/RNG.h:21-42
 * ---------------------------------------------------------------------------------------- */
#define WELL_R 624
#define WELL_M1 70
#define WELL_M2 179
#define WELL_M3 449
#define WELL_MASKU 0x7fffffffU /* 0xffffffff >> (32 - 31) */
#define WELL_MASKL 0x80000000U
#define WELL_TEMPER 0x41180000U
#define WELL_BURNIN 500

typedef struct {
	uint32_t state[WELL_R];
	int idx;
} well_t;

static uint32_t well_next(well_t* g) {
	const int i = g->idx;
	uint32_t* s = g->state;
	/* The reference's "Under" operands (stateIndex 0 and 1) read STATE[i + M3 - 1] and
	 * STATE[i + M3 - 2] rather than wrapping to R-1 / R-2 (src/RNG.cpp:26,28 used at :93,:106).
	 * Restated literally: this is what the package's streams are. */
	const uint32_t vrm1 = (i >= 1) ? s[i - 1] : s[i + WELL_M3 - 1];
	const uint32_t vrm2 = (i >= 2) ? s[i - 2] : s[i + WELL_M3 - 2];
	const uint32_t v0 = s[i];
	const uint32_t vm1 = s[(i + WELL_M1) % WELL_R];
	const uint32_t vm2 = s[(i + WELL_M2) % WELL_R];
	const uint32_t vm3 = s[(i + WELL_M3) % WELL_R];
	const uint32_t z0 = (vrm1 & WELL_MASKL) | (vrm2 & WELL_MASKU);
	const uint32_t z1 = (v0 ^ (v0 << 25)) ^ (vm1 ^ (vm1 >> 27));
	const uint32_t z2 = (vm2 >> 9) ^ (vm3 ^ (vm3 >> 1));
	const uint32_t nv1 = z1 ^ z2;
	const int ni = (i + WELL_R - 1) % WELL_R;
	s[i] = nv1;
	s[ni] = z0 ^ (z1 ^ (z1 << 9)) ^ (z2 ^ (z2 << 21)) ^ (nv1 ^ (nv1 >> 21));
	g->idx = ni;
	return s[ni] ^ (s[(ni + WELL_M2 + 1) % WELL_R] & WELL_TEMPER);
}

static void well_burn(well_t* g) {
	for (int i = 0; i < WELL_BURNIN; ++i) (void) well_next(g);
}

/* src/RNG.cpp:75-90 */
static void well_seed_u32(well_t* g, uint32_t seed) {
	g->state[0] = seed;
	for (uint32_t i = 1; i < WELL_R; ++i) {
		g->state[i] = 1812433253U * (g->state[i - 1] ^ (g->state[i - 1] >> 30)) + i;
	}
	g->idx = 0;
	well_burn(g);
}

/* src/RNG.cpp:61-73 */
static void well_seed_vec(well_t* g, const uint32_t* seedvec) {
	memcpy(g->state, seedvec, sizeof(uint32_t) * WELL_R);
	g->idx = 0;
	well_burn(g);
}

/* src/RNG.h:21-23 */
static double well_uniform(well_t* g, double lo, double hi) {
	return lo + ((double) well_next(g) / 4294967296.0) * (hi - lo);
}

/* src/ShuffledSet.cpp:38-44: in-place pass over a persistent permutation */
static void shuffle_all(uint32_t* set, int n, well_t* g) {
	for (int i = 0; i < n; ++i) {
		const int j = (int) (uint64_t) well_uniform(g, (double) i, (double) n);
		const uint32_t tmp = set[i];
		set[i] = set[j];
		set[j] = tmp;
	}
}

static int cmp_u32(const void* a, const void* b) {
	const uint32_t x = *(const uint32_t*) a, y = *(const uint32_t*) b;
	return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------
 * Evaluator object
 * ---------------------------------------------------------------------------------------- */
enum { KIND_PLS = 0, KIND_FIT = 1, KIND_LM = 2 };
enum { STAT_BIC = 0, STAT_AIC = 1, STAT_ADJ_R2 = 2, STAT_R2 = 3 };

typedef struct {
	int kind;
	int n, p;
	double* X; /* column-major n x p */
	double* y;
	int reps, outer, inner; /* PLS: R, O, I(effective); FIT: inner = K */
	int maxNComp;
	double sdfact; /* already divided by sqrt(I) */
	double gamma;
	int stat;
	double r2denom;
	int nseg;
	uint32_t* seg_rows;
	uint32_t* seg_off;
} evaluator_t;

typedef struct {
	uint32_t* rows;
	uint32_t* off;
	int count, cap_count;
	long long total, cap_total;
} seglist_t;

static void seg_push(seglist_t* L, const uint32_t* src, int len, int sorted_copy) {
	if (L->count + 1 >= L->cap_count) {
		L->cap_count = L->cap_count ? 2 * L->cap_count : 64;
		L->off = (uint32_t*) realloc(L->off, sizeof(uint32_t) * (size_t) (L->cap_count + 1));
	}
	if (L->total + len > L->cap_total) {
		L->cap_total = 2 * (L->total + len) + 1024;
		L->rows = (uint32_t*) realloc(L->rows, sizeof(uint32_t) * (size_t) L->cap_total);
	}
	memcpy(L->rows + L->total, src, sizeof(uint32_t) * (size_t) len);
	if (sorted_copy) qsort(L->rows + L->total, (size_t) len, sizeof(uint32_t), cmp_u32);
	L->off[L->count] = (uint32_t) L->total;
	L->total += len;
	L->count += 1;
	L->off[L->count] = (uint32_t) L->total;
}

/* src/PLSEvaluator.cpp:27-57 (constructor) and :96-223 (initSegmentation) */
static int pls_init_segmentation(evaluator_t* e, const uint32_t* seedvec, int innerIn, int outerIn, double testSetSize,
                                 double sdfactIn) {
	const int n = e->n;
	e->outer = (outerIn < 1) ? 1 : outerIn;                                       /* :33 */
	e->inner = (e->outer <= 1 && testSetSize == 0.0) ? innerIn - 1 : innerIn;      /* :34 */
	if (e->inner < 1) return 1;
	e->sdfact = sdfactIn / sqrt((double) e->inner);                                /* :35 */
	if (e->outer > 1) testSetSize = 1.0 / (double) e->outer;                       /* :46-48 */
	if (testSetSize < 0.0 || testSetSize >= 1.0) return 1;                         /* :50-52 */
	if (testSetSize == 0.0 && e->outer == 1) testSetSize = 1.0 / (e->inner + 1.0); /* :101-103 */

	const uint64_t olenBase = (uint64_t) ((double) n * testSetSize);               /* :109 double product, truncated */
	if (olenBase == 0) return 1; /* the reference divides by zero here (n % 0) */
	const uint64_t oremInit = (uint64_t) n % olenBase;                             /* :110 */
	const uint64_t ilenBase = ((uint64_t) n - olenBase) / (uint64_t) e->inner;     /* :121 */
	const uint64_t iremBase = ((uint64_t) n - olenBase) % (uint64_t) e->inner;     /* :122 */
	uint64_t ilenBig = ilenBase;
	if (oremInit > 0 && iremBase == 0) ilenBig -= 1;                               /* :125-127 */

	const uint64_t minFit = (uint64_t) n - olenBase - ilenBase - 2;                /* :135 */
	if (minFit <= (uint64_t) e->maxNComp || e->maxNComp == 0) e->maxNComp = (int) minFit; /* :136-138 */

	well_t g;
	well_seed_vec(&g, seedvec);                                                    /* :97 */
	uint32_t* set = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) n);             /* :98, never reset */
	uint32_t* sh = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) n);
	uint32_t* tmp = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) n);
	for (int i = 0; i < n; ++i) set[i] = (uint32_t) i;
	seglist_t L;
	memset(&L, 0, sizeof(L));

	for (int rep = 0; rep < e->reps; ++rep) {
		shuffle_all(set, n, &g);                                                   /* :145 */
		memcpy(sh, set, sizeof(uint32_t) * (size_t) n);
		uint64_t orem = oremInit, irem;
		if (orem > 0) irem = (iremBase == 0) ? (uint64_t) e->inner - 1 : iremBase - 1; /* :147-153 */
		else irem = iremBase;                                                      /* :155 */
		uint64_t olen = 0;
		for (int o = 0; o < e->outer; ++o) {
			uint64_t ilenFixed = ilenBase;
			if (orem > 0) { olen = olenBase + 1; ilenFixed = ilenBig; --orem; }      /* :161-164 */
			else { olen = olenBase; irem = iremBase; }                              /* :165-168 */
			uint64_t pos = 0;
			for (int j = 0; j < e->inner; ++j) {
				const uint64_t ilen = ilenFixed + (((uint64_t) j < irem) ? 1 : 0);   /* :174-178 */
				const uint64_t ntr = (uint64_t) n - olen - ilen;
				/* training rows: sh[0,pos) followed by sh[pos+ilen, n-olen)  (:183-193) */
				memcpy(tmp, sh, sizeof(uint32_t) * pos);
				memcpy(tmp + pos, sh + pos + ilen, sizeof(uint32_t) * (ntr - pos));
				seg_push(&L, tmp, (int) ntr, 1);                                     /* :193-196 */
				seg_push(&L, sh + pos, (int) ilen, 1);                               /* :199 */
				pos += ilen;
			}
			seg_push(&L, sh, (int) pos, 1);                                          /* :210 outer training */
			seg_push(&L, sh + pos, (int) ((uint64_t) n - pos), 1);                   /* :213 outer test */
			/* rotate: last block to the front (:220) */
			memcpy(tmp, sh + pos, sizeof(uint32_t) * ((uint64_t) n - pos));
			memcpy(tmp + ((uint64_t) n - pos), sh, sizeof(uint32_t) * pos);
			memcpy(sh, tmp, sizeof(uint32_t) * (size_t) n);
		}
	}
	free(set); free(sh); free(tmp);
	e->nseg = L.count;
	e->seg_rows = L.rows;
	e->seg_off = L.off;
	return 0;
}

/* src/BICEvaluator.cpp:22-44 (constructor) and :96-139 (initSegmentation) */
static int fit_init_segmentation(evaluator_t* e, const uint32_t* seedvec, int numSegments, double sdfactIn) {
	const int n = e->n;
	if (numSegments < 2) return 1;                                                 /* :31-33 */
	e->inner = numSegments;
	e->sdfact = sdfactIn / sqrt((double) numSegments);                             /* :26 */
	/* r2denom = n * population variance of y (:37) */
	double m = 0.0;
	for (int i = 0; i < n; ++i) m += e->y[i];
	m /= (double) n;
	double ss = 0.0;
	for (int i = 0; i < n; ++i) ss += (e->y[i] - m) * (e->y[i] - m);
	e->r2denom = (double) n * (ss / (double) n);
	if (e->maxNComp <= 1) e->maxNComp = n - 1;                                     /* :39-41 */

	well_t g;
	well_seed_vec(&g, seedvec);
	uint32_t* sh = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) n);
	uint32_t* tmp = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) n);
	for (int i = 0; i < n; ++i) sh[i] = (uint32_t) i;
	shuffle_all(sh, n, &g);                                                        /* :98 */
	const int len = n / numSegments, rem = n % numSegments;                        /* :99-100 */
	if (n - len - 2 < e->maxNComp) e->maxNComp = n - len - 2;                      /* :106-108 */
	seglist_t L;
	memset(&L, 0, sizeof(L));
	int pos = 0;
	for (int i = 0; i < numSegments; ++i) {
		const int sl = len + ((i < rem) ? 1 : 0);                                  /* :113-116 */
		memcpy(tmp, sh, sizeof(uint32_t) * (size_t) pos);                          /* :120-126 */
		memcpy(tmp + pos, sh + pos + sl, sizeof(uint32_t) * (size_t) (n - sl - pos));
		seg_push(&L, tmp, n - sl, 1);
		seg_push(&L, sh + pos, sl, 1);                                             /* :134 */
		pos += sl;
	}
	free(sh); free(tmp);
	e->nseg = L.count;
	e->seg_rows = L.rows;
	e->seg_off = L.off;
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * SIMPLS in X-space, the reference's operation order
 * follows src/PLSSimpls.cpp:28-49 (centerView), :57-100 (helpers), :106-166 (fit)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	int cap_n, cap_k, cap_a;
	double* Xsel;  /* n x k: columns of the chromosome (PLS::viewSelectColumns, src/PLS.cpp:13-17) */
	double* Xtr;   /* ntr x k: row view (src/PLS.cpp:19-23), centred in place */
	double* ytr;
	double* xm;
	double* S;
	double* t;
	double* V;     /* k x A */
	double* coef;  /* k x A, cumulative columns */
	double* b0;
	double* msep_mean; double* msep_m2; /* Welford per component (src/OnlineStddev.h) */
	double* pred;
} scratch_t;

static void scratch_init(scratch_t* s, int n, int k, int a) {
	s->cap_n = n; s->cap_k = k; s->cap_a = a;
	s->Xsel = (double*) malloc(sizeof(double) * (size_t) n * (size_t) k);
	s->Xtr = (double*) malloc(sizeof(double) * (size_t) n * (size_t) k);
	s->ytr = (double*) malloc(sizeof(double) * (size_t) n);
	s->xm = (double*) malloc(sizeof(double) * (size_t) k);
	s->S = (double*) malloc(sizeof(double) * (size_t) k);
	s->t = (double*) malloc(sizeof(double) * (size_t) n);
	s->V = (double*) malloc(sizeof(double) * (size_t) k * (size_t) a);
	s->coef = (double*) malloc(sizeof(double) * (size_t) k * (size_t) a);
	s->b0 = (double*) malloc(sizeof(double) * (size_t) a);
	s->msep_mean = (double*) malloc(sizeof(double) * (size_t) a);
	s->msep_m2 = (double*) malloc(sizeof(double) * (size_t) a);
	s->pred = (double*) malloc(sizeof(double) * (size_t) n);
}

static void scratch_free(scratch_t* s) {
	free(s->Xsel); free(s->Xtr); free(s->ytr); free(s->xm); free(s->S); free(s->t); free(s->V);
	free(s->coef); free(s->b0); free(s->msep_mean); free(s->msep_m2); free(s->pred);
}

static double dotp(const double* restrict a, const double* restrict b, int n) {
	double s = 0.0;
#pragma omp simd reduction(+ : s)
	for (int i = 0; i < n; ++i) s += a[i] * b[i];
	return s;
}

/* Fit on rows `rows[0..ntr)` (NULL = all n rows) of the selected columns; ncomp >= 1.
 * Returns 0, or 2 on the underflow of src/PLSSimpls.cpp:141-143. */
static int simpls_fit(const evaluator_t* e, scratch_t* s, int k, const uint32_t* rows, int ntr, int ncomp) {
	const int n = e->n;
	double* restrict Xt = s->Xtr;
	/* row view + centring (src/PLS.cpp:19-23, src/PLSSimpls.cpp:44-48) */
	double ym = 0.0;
	for (int i = 0; i < ntr; ++i) { s->ytr[i] = e->y[rows ? rows[i] : (uint32_t) i]; ym += s->ytr[i]; }
	ym /= (double) ntr;
	for (int i = 0; i < ntr; ++i) s->ytr[i] -= ym;
	for (int j = 0; j < k; ++j) {
		const double* restrict src = s->Xsel + (size_t) j * (size_t) n;
		double* restrict dst = Xt + (size_t) j * (size_t) ntr;
		double m = 0.0;
		for (int i = 0; i < ntr; ++i) { dst[i] = src[rows ? rows[i] : (uint32_t) i]; m += dst[i]; }
		m /= (double) ntr;
		for (int i = 0; i < ntr; ++i) dst[i] -= m;
		s->xm[j] = m;
	}
	/* S = X'y (:126) */
	for (int j = 0; j < k; ++j) s->S[j] = dotp(Xt + (size_t) j * (size_t) ntr, s->ytr, ntr);

	for (int a = 0; a < ncomp; ++a) {
		double* restrict v = s->V + (size_t) a * (size_t) k;
		double* restrict t = s->t;
		/* t = X S, centred, normalised (:135-145) */
		for (int i = 0; i < ntr; ++i) t[i] = 0.0;
		for (int j = 0; j < k; ++j) {
			const double sj = s->S[j];
			const double* restrict col = Xt + (size_t) j * (size_t) ntr;
			for (int i = 0; i < ntr; ++i) t[i] += col[i] * sj;
		}
		double tm = 0.0;
		for (int i = 0; i < ntr; ++i) tm += t[i];
		tm /= (double) ntr;
		for (int i = 0; i < ntr; ++i) t[i] -= tm;
		const double tnorm = sqrt(dotp(t, t, ntr));
		if (tnorm < 1e-25) return 2;                                               /* :141-143, NORM_TOL :16 */
		for (int i = 0; i < ntr; ++i) t[i] /= tnorm;
		/* v = X't, q = y't (:147-148) */
		for (int j = 0; j < k; ++j) v[j] = dotp(Xt + (size_t) j * (size_t) ntr, t, ntr);
		const double q = dotp(s->ytr, t, ntr);
		/* modified Gram-Schmidt against the previous loadings (:151, :57-63) */
		for (int b = 0; b < a; ++b) {
			const double* restrict vb = s->V + (size_t) b * (size_t) k;
			const double d = dotp(vb, v, k);
			for (int j = 0; j < k; ++j) v[j] -= vb[j] * d;
		}
		/* cumulative coefficients from S BEFORE deflation (:152-156, :95-100) */
		double* restrict ca = s->coef + (size_t) a * (size_t) k;
		if (a > 0) {
			const double* restrict cp = s->coef + (size_t) (a - 1) * (size_t) k;
			const double f = q / tnorm;
			for (int j = 0; j < k; ++j) ca[j] = cp[j] + s->S[j] * f;
		} else {
			for (int j = 0; j < k; ++j) ca[j] = s->S[j] * q / tnorm;
		}
		/* deflate S and normalise v (:160, :70-88) */
		double nrm = 0.0, d = 0.0;
		for (int j = 0; j < k; ++j) { nrm += v[j] * v[j]; d += v[j] * s->S[j]; }
		d /= nrm;
		nrm = sqrt(nrm);
		for (int j = 0; j < k; ++j) { s->S[j] -= v[j] * d; v[j] /= nrm; }
		/* intercept (:162) */
		s->b0[a] = ym - dotp(s->xm, ca, k);
	}
	return 0;
}

/* prediction on uncentred rows with `a1` (1-based) components: src/PLS.cpp:34-47 */
static void simpls_predict(const evaluator_t* e, scratch_t* s, int k, const uint32_t* rows, int nte, int a1) {
	const int n = e->n;
	const double* restrict c = s->coef + (size_t) (a1 - 1) * (size_t) k;
	double* restrict pr = s->pred;
	for (int i = 0; i < nte; ++i) pr[i] = 0.0;
	for (int j = 0; j < k; ++j) {
		const double* restrict col = s->Xsel + (size_t) j * (size_t) n;
		const double cj = c[j];
		for (int i = 0; i < nte; ++i) pr[i] += col[rows ? rows[i] : (uint32_t) i] * cj;
	}
	for (int i = 0; i < nte; ++i) pr[i] += s->b0[a1 - 1];
}

/* Welford update: src/OnlineStddev.h:36-40 */
static void welford(double* mean, double* m2, uint16_t* cnt, double x) {
	const double delta = x - *mean;
	*cnt = (uint16_t) (*cnt + 1);
	*mean += delta / (double) *cnt;
	*m2 += delta * (x - *mean);
}

/* K-fold MSEP table + one-SE rule; segments [first, first + 2*folds) are (train,test) pairs.
 * follows src/PLSEvaluator.cpp:255-308 (identical in src/BICEvaluator.cpp:162-210).
 * Returns 0 and the 1-based optimal component count, or 2 on underflow. */
static int cv_select_ncomp(const evaluator_t* e, scratch_t* s, int k, int A, int first, int folds, int* optOut) {
	for (int a = 0; a < A; ++a) { s->msep_mean[a] = 0.0; s->msep_m2[a] = 0.0; }
	for (int f = 0; f < folds; ++f) {
		const int tr = first + 2 * f, te = tr + 1;
		const uint32_t* trRows = e->seg_rows + e->seg_off[tr];
		const int ntr = (int) (e->seg_off[tr + 1] - e->seg_off[tr]);
		const uint32_t* teRows = e->seg_rows + e->seg_off[te];
		const int nte = (int) (e->seg_off[te + 1] - e->seg_off[te]);
		const int rc = simpls_fit(e, s, k, trRows, ntr, A);
		if (rc) return rc;
		for (int a = 0; a < A; ++a) {
			simpls_predict(e, s, k, teRows, nte, a + 1);
			double mse = 0.0;
			for (int i = 0; i < nte; ++i) { const double r = e->y[teRows[i]] - s->pred[i]; mse += r * r; }
			mse /= (double) nte;
			uint16_t c = (uint16_t) f; /* every component has seen f folds so far */
			welford(&s->msep_mean[a], &s->msep_m2[a], &c, mse);
		}
	}
	const int cnt = folds;
	/* first minimum (:284-292) */
	int minIdx = 0;
	double cutoff = s->msep_mean[0];
	for (int a = 1; a < A; ++a) {
		if (s->msep_mean[a] < cutoff) { minIdx = a; cutoff = s->msep_mean[a]; }
	}
	/* + sd * sdfact (:296); sd = sqrt(M2 / (cnt - 1)), NaN when folds == 1 */
	cutoff += sqrt(s->msep_m2[minIdx] / (double) ((int) cnt - 1)) * e->sdfact;
	int opt;
	if (minIdx == 0) {
		opt = 1;                                                                   /* :298-299 */
	} else {
		opt = 0;                                                                   /* :301-307 */
		while (opt < minIdx && s->msep_mean[opt] > cutoff) ++opt;
		if (opt <= minIdx) ++opt;
	}
	*optOut = opt;
	return 0;
}

static void gather_columns(const evaluator_t* e, scratch_t* s, const uint32_t* cols, int k) {
	for (int j = 0; j < k; ++j) {
		memcpy(s->Xsel + (size_t) j * (size_t) e->n, e->X + (size_t) cols[j] * (size_t) e->n, sizeof(double) * (size_t) e->n);
	}
}

/* src/PLSEvaluator.cpp:71-91 (evaluate) and :228-343 (estSEP) */
static int pls_evaluate(const evaluator_t* e, const uint32_t* cols, int k, double* fitness) {
	if (k == 0) return 1;                                                          /* :72-75 */
	const int A = (e->maxNComp < k) ? e->maxNComp : k;                             /* :81 */
	scratch_t s;
	scratch_init(&s, e->n, k, A > 0 ? A : 1);
	gather_columns(e, &s, cols, k);                                                /* :79 */
	double sumSEP = 0.0;
	int seg = 0, rc = 0;
	for (int rep = 0; rep < e->reps && !rc; ++rep) {
		double pm = 0.0, pm2 = 0.0;
		uint16_t pcnt = 0;                                                         /* predSD.reset() :245 */
		for (int o = 0; o < e->outer && !rc; ++o) {
			int opt = 1;
			rc = cv_select_ncomp(e, &s, k, A, seg, e->inner, &opt);                  /* :255-308 */
			if (rc) break;
			seg += 2 * e->inner;
			const uint32_t* trRows = e->seg_rows + e->seg_off[seg];
			const int ntr = (int) (e->seg_off[seg + 1] - e->seg_off[seg]);
			const uint32_t* teRows = e->seg_rows + e->seg_off[seg + 1];
			const int nte = (int) (e->seg_off[seg + 2] - e->seg_off[seg + 1]);
			rc = simpls_fit(e, &s, k, trRows, ntr, opt);                             /* :316-317 */
			if (rc) break;
			simpls_predict(e, &s, k, teRows, nte, opt);
			for (int i = 0; i < nte; ++i) welford(&pm, &pm2, &pcnt, e->y[teRows[i]] - s.pred[i]); /* :323-325 */
			seg += 2;
		}
		sumSEP += sqrt(pm2 / (double) ((int) pcnt - 1));                            /* :335 */
	}
	scratch_free(&s);
	if (rc) return rc;                                                             /* :337-340 */
	*fitness = sumSEP * (-(1.0 + e->gamma * (double) k) / (double) e->reps);       /* :87 */
	return 0;
}

/* src/BICEvaluator.cpp:54-91 (evaluate) and :144-233 (getRSS) */
static int fit_evaluate(const evaluator_t* e, const uint32_t* cols, int k, double* fitness) {
	if (k == 0) return 1;                                                          /* :55-58 */
	const int A = (e->maxNComp < k) ? e->maxNComp : k;                             /* :63 */
	const int n = e->n;
	scratch_t s;
	scratch_init(&s, n, k, A > 0 ? A : 1);
	gather_columns(e, &s, cols, k);
	int opt = 1;
	int rc = cv_select_ncomp(e, &s, k, A, 0, e->inner, &opt);
	double RSS = 0.0;
	if (!rc) rc = simpls_fit(e, &s, k, NULL, n, opt);                               /* :223-224 all rows */
	if (!rc) {
		simpls_predict(e, &s, k, NULL, n, opt);
		for (int i = 0; i < n; ++i) { const double r = e->y[i] - s.pred[i]; RSS += r * r; } /* :226 */
	}
	scratch_free(&s);
	if (rc) return rc;
	const double df = (double) k + 2.0;                                            /* :64 */
	const double dn = (double) n;
	switch (e->stat) {
	case STAT_BIC: *fitness = -(dn * log(RSS / dn) + df * log(dn)); break;         /* :70 */
	case STAT_AIC: *fitness = -(dn * log(RSS / dn) + df * 2.0); break;             /* :74 */
	case STAT_ADJ_R2: {                                                            /* :78-79 */
		const double r2 = 1.0 - (RSS / e->r2denom);
		*fitness = 1.0 - (((dn - 1.0) * (1.0 - r2)) / (dn - df - 1.0));
		break;
	}
	default: *fitness = 1.0 - (RSS / e->r2denom); break;                           /* :83 */
	}
	return 0;
}

/* src/LMEvaluator.cpp:27-67; arma::solve restated as Householder QR least squares */
static int lm_evaluate(const evaluator_t* e, const uint32_t* cols, int k, double* fitness) {
	const int n = e->n, m = k + 1;                                                 /* intercept column first, :29-32 */
	if (n < m) return 3;
	double* A = (double*) malloc(sizeof(double) * (size_t) n * (size_t) m);
	double* r = (double*) malloc(sizeof(double) * (size_t) n);
	double* rd = (double*) malloc(sizeof(double) * (size_t) m);
	double* b = (double*) malloc(sizeof(double) * (size_t) m);
	for (int i = 0; i < n; ++i) { A[i] = 1.0; r[i] = e->y[i]; }
	for (int j = 0; j < k; ++j) memcpy(A + (size_t) (j + 1) * (size_t) n, e->X + (size_t) cols[j] * (size_t) n, sizeof(double) * (size_t) n);
	int rc = 0;
	double maxd = 0.0;
	for (int c = 0; c < m && !rc; ++c) {
		double* restrict ac = A + (size_t) c * (size_t) n;
		double nrm = sqrt(dotp(ac + c, ac + c, n - c));
		if (nrm == 0.0) { rc = 3; break; }
		const double alpha = (ac[c] > 0.0) ? -nrm : nrm;
		ac[c] -= alpha;
		const double vtv = dotp(ac + c, ac + c, n - c);
		if (vtv > 0.0) {
			for (int j = c + 1; j < m; ++j) {
				double* restrict aj = A + (size_t) j * (size_t) n;
				const double f = 2.0 * dotp(ac + c, aj + c, n - c) / vtv;
				for (int i = c; i < n; ++i) aj[i] -= f * ac[i];
			}
			const double f = 2.0 * dotp(ac + c, r + c, n - c) / vtv;
			for (int i = c; i < n; ++i) r[i] -= f * ac[i];
		}
		rd[c] = alpha;
		if (fabs(alpha) > maxd) maxd = fabs(alpha);
	}
	for (int c = 0; c < m && !rc; ++c) if (fabs(rd[c]) <= maxd * 1e-13 * (double) n) rc = 3; /* :60-61 singular */
	double RSS = 0.0;
	if (!rc) {
		for (int c = m - 1; c >= 0; --c) {
			double v = r[c];
			for (int j = c + 1; j < m; ++j) v -= A[(size_t) j * (size_t) n + (size_t) c] * b[j];
			b[c] = v / rd[c];
		}
		/* explicit residuals y - Xsub b on the ORIGINAL columns (:35-37) */
		for (int i = 0; i < n; ++i) r[i] = e->y[i] - b[0];
		for (int j = 0; j < k; ++j) {
			const double* restrict col = e->X + (size_t) cols[j] * (size_t) n;
			const double bj = b[j + 1];
			for (int i = 0; i < n; ++i) r[i] -= col[i] * bj;
		}
		RSS = dotp(r, r, n);
	}
	free(A); free(r); free(rd); free(b);
	if (rc) return rc;
	const double dn = (double) n, dm = (double) m;
	switch (e->stat) {
	case STAT_BIC: *fitness = -(dn * log(RSS / dn) + dm * log(dn)); break;         /* :41 */
	case STAT_AIC: *fitness = -(dn * log(RSS / dn) + dm * 2.0); break;             /* :45 */
	case STAT_ADJ_R2: {                                                            /* :49-50 */
		const double r2 = 1.0 - (RSS / e->r2denom);
		*fitness = 1.0 - (((dn - 1.0) * (1.0 - r2)) / (dn - dm - 1.0));
		break;
	}
	default: *fitness = 1.0 - (RSS / e->r2denom); break;                           /* :54 */
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * C entry points (same shapes as oracle/ref_shim.cpp)
 * ---------------------------------------------------------------------------------------- */
void orc_rng_u32_stream(uint32_t seed, int n, uint32_t* out) {
	well_t g;
	well_seed_u32(&g, seed);
	for (int i = 0; i < n; ++i) out[i] = well_next(&g);
}

/* two-stage seeding of src/GenAlg.cpp:99-103 */
void orc_seedvec(uint32_t seed, uint32_t* out624) {
	orc_rng_u32_stream(seed, WELL_R, out624);
}

void orc_rng_vec_stream(const uint32_t* seedvec, int n, uint32_t* out) {
	well_t g;
	well_seed_vec(&g, seedvec);
	for (int i = 0; i < n; ++i) out[i] = well_next(&g);
}

void orc_rng_vec_uniform(const uint32_t* seedvec, double lo, double hi, int n, double* out) {
	well_t g;
	well_seed_vec(&g, seedvec);
	for (int i = 0; i < n; ++i) out[i] = well_uniform(&g, lo, hi);
}

void orc_shuffle_all(const uint32_t* seedvec, int size, int ncalls, uint32_t* out) {
	well_t g;
	well_seed_vec(&g, seedvec);
	uint32_t* set = (uint32_t*) malloc(sizeof(uint32_t) * (size_t) size);
	for (int i = 0; i < size; ++i) set[i] = (uint32_t) i;
	for (int c = 0; c < ncalls; ++c) {
		shuffle_all(set, size, &g);
		memcpy(out + (size_t) c * (size_t) size, set, sizeof(uint32_t) * (size_t) size);
	}
	free(set);
}

static evaluator_t* evaluator_new(int kind, const double* X, const double* y, int n, int p) {
	evaluator_t* e = (evaluator_t*) calloc(1, sizeof(evaluator_t));
	e->kind = kind; e->n = n; e->p = p;
	e->X = (double*) malloc(sizeof(double) * (size_t) n * (size_t) p);
	e->y = (double*) malloc(sizeof(double) * (size_t) n);
	memcpy(e->X, X, sizeof(double) * (size_t) n * (size_t) p);
	memcpy(e->y, y, sizeof(double) * (size_t) n);
	return e;
}

void orc_evaluator_destroy(void* h) {
	evaluator_t* e = (evaluator_t*) h;
	if (!e) return;
	free(e->X); free(e->y); free(e->seg_rows); free(e->seg_off); free(e);
}

void* orc_pls_evaluator_create(const double* X, const double* y, int n, int p, int numReplications, int maxNComp,
                               const uint32_t* seedvec, int innerSegments, int outerSegments, double testSetSize,
                               double sdfact, double complexityPenalty) {
	evaluator_t* e = evaluator_new(KIND_PLS, X, y, n, p);
	e->reps = numReplications; e->maxNComp = maxNComp; e->gamma = complexityPenalty;
	if (pls_init_segmentation(e, seedvec, innerSegments, outerSegments, testSetSize, sdfact)) {
		orc_evaluator_destroy(e);
		return NULL;
	}
	return e;
}

void* orc_fit_evaluator_create(const double* X, const double* y, int n, int p, int maxNComp, const uint32_t* seedvec,
                               int numSegments, int statistic, double sdfact) {
	evaluator_t* e = evaluator_new(KIND_FIT, X, y, n, p);
	e->maxNComp = maxNComp; e->stat = statistic;
	if (fit_init_segmentation(e, seedvec, numSegments, sdfact)) {
		orc_evaluator_destroy(e);
		return NULL;
	}
	return e;
}

void* orc_lm_evaluator_create(const double* X, const double* y, int n, int p, int statistic) {
	evaluator_t* e = evaluator_new(KIND_LM, X, y, n, p);
	e->stat = statistic;
	/* r2denom = sum (y - mean y)^2 (src/LMEvaluator.cpp:24) */
	double m = 0.0;
	for (int i = 0; i < n; ++i) m += y[i];
	m /= (double) n;
	double ss = 0.0;
	for (int i = 0; i < n; ++i) ss += (y[i] - m) * (y[i] - m);
	e->r2denom = ss;
	return e;
}

int orc_segmentation_count(void* h) { return ((evaluator_t*) h)->nseg; }

long long orc_segmentation_total(void* h) {
	evaluator_t* e = (evaluator_t*) h;
	return e->nseg ? (long long) e->seg_off[e->nseg] : 0;
}

void orc_segmentation_get(void* h, uint32_t* rows, uint32_t* offsets) {
	evaluator_t* e = (evaluator_t*) h;
	if (!e->nseg) { offsets[0] = 0; return; }
	memcpy(rows, e->seg_rows, sizeof(uint32_t) * (size_t) e->seg_off[e->nseg]);
	memcpy(offsets, e->seg_off, sizeof(uint32_t) * (size_t) (e->nseg + 1));
}

typedef struct {
	const evaluator_t* e;
	const uint32_t* idx;
	const uint32_t* offsets;
	int begin, end;
	double* fitness;
	int32_t* status;
} slice_t;

static void* slice_run(void* arg) {
	slice_t* s = (slice_t*) arg;
	for (int c = s->begin; c < s->end; ++c) {
		const uint32_t* cols = s->idx + s->offsets[c];
		const int k = (int) (s->offsets[c + 1] - s->offsets[c]);
		double f = 0.0;
		int rc;
		switch (s->e->kind) {
		case KIND_PLS: rc = pls_evaluate(s->e, cols, k, &f); break;
		case KIND_FIT: rc = fit_evaluate(s->e, cols, k, &f); break;
		default: rc = lm_evaluate(s->e, cols, k, &f); break;
		}
		s->fitness[c] = rc ? 0.0 : f;
		s->status[c] = rc;
	}
	return NULL;
}

/* batch loop of src/GenAlg.cpp:312-325 over contiguous slices, one host thread each
 * (the slicing rule of src/MultiThreadedPopulation.cpp:260-262,302-308) */
int orc_evaluate(void* h, const uint32_t* idx, const uint32_t* offsets, int nsub, int nthreads, double* fitness,
                 int32_t* status, long long* elapsed_ns) {
	const evaluator_t* e = (const evaluator_t*) h;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > nsub) nthreads = nsub > 0 ? nsub : 1;
	slice_t* sl = (slice_t*) malloc(sizeof(slice_t) * (size_t) nthreads);
	pthread_t* th = (pthread_t*) malloc(sizeof(pthread_t) * (size_t) nthreads);
	const int per = nsub / nthreads, rem = nsub % nthreads;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < nthreads; ++t) {
		sl[t].e = e; sl[t].idx = idx; sl[t].offsets = offsets; sl[t].fitness = fitness; sl[t].status = status;
		sl[t].begin = t * per + (t < rem ? t : rem);
		sl[t].end = sl[t].begin + per + (t < rem ? 1 : 0);
		if (t > 0) pthread_create(&th[t], NULL, slice_run, &sl[t]);
	}
	slice_run(&sl[0]);
	for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (elapsed_ns) *elapsed_ns = (long long) (t1.tv_sec - t0.tv_sec) * 1000000000LL + (long long) (t1.tv_nsec - t0.tv_nsec);
	free(sl); free(th);
	return 0;
}

/* raw SIMPLS (the hidden C_simpls entry, src/GenAlg.cpp:351-371, plus a row view) */
int orc_simpls(const double* X, const double* y, int n, int p, const uint32_t* cols, int k, const uint32_t* rows,
               int nrows, int ncomp, double* coef, double* b0) {
	evaluator_t* e = evaluator_new(KIND_PLS, X, y, n, p);
	scratch_t s;
	scratch_init(&s, n, k, ncomp);
	gather_columns(e, &s, cols, k);
	const int rc = simpls_fit(e, &s, k, nrows > 0 ? rows : NULL, nrows > 0 ? nrows : n, ncomp);
	if (!rc) {
		memcpy(coef, s.coef, sizeof(double) * (size_t) k * (size_t) ncomp);
		memcpy(b0, s.b0, sizeof(double) * (size_t) ncomp);
	}
	scratch_free(&s);
	orc_evaluator_destroy(e);
	return rc;
}

int orc_hardware_threads(void) { return (int) sysconf(_SC_NPROCESSORS_ONLN); }
