/*
 * cutesv_oracle.c -- CPU restatement of cuteSV's sort -> cluster -> consensus -> genotype path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cutesv_b200/ may import, link or execute this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference ITSELF: oracle/gen_golden.py imports the
 * unmodified reference from /root/reference/src (pysam stubbed), runs resolution_DEL/INS/INV/
 * DUP/TRA and cal_GL on seeded inputs and commits inputs + outputs under tests/golden/;
 * tests/test_oracle_golden.py replays them through this file.
 *
 * Every function cites the reference lines it follows ("cuteSV:N" = src/cuteSV/cuteSV).
 * Deliberately sequential and literal: floating point is evaluated in the same order as the
 * Python/numpy code (compile with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/cutesv_b200.h"

/* ------------------------------------------------------------------------------------------ */
/* cal_GL and friends: cuteSV_genotype.py:10-60                                               */
/* ------------------------------------------------------------------------------------------ */

static double o_rint(double x) { return nearbyint(x); } /* np.around / round(): half-to-even */

/* rescale_read_counts, cuteSV_genotype.py:25-31 */
static void rescale_read_counts(int* c0, int* c1) {
    int total = *c0 + *c1;
    if (total > 100) {
        double f = (double)*c0 / (double)total;
        *c0 = (int)(100.0 * f);
        *c1 = 100 - *c0;
    }
}

/* cal_GL, cuteSV_genotype.py:33-56 (with log10sumexp :14-17 and normalize_log10_probs :19-23) */
void csvo_cal_gl(int c0, int c1, csv_geno* g) {
    g->dr = c0;
    g->dv = c1;
    g->status = 0;
    if (c0 == 3 && c1 == 1) {
        g->gt = 1; g->pl[0] = 3; g->pl[1] = 3; g->pl[2] = 24; g->gq = 3; g->qual = 3.0;
        return;
    }
    if (c0 == 6 && c1 == 2) {
        g->gt = 1; g->pl[0] = 3; g->pl[1] = 3; g->pl[2] = 45; g->gq = 3; g->qual = 3.0;
        return;
    }
    rescale_read_counts(&c0, &c1);
    const double err = 0.1;
    const double prior = (double)(1.0 / 3.0);
    double gl00 = pow((1 - err), (double)c0) * pow(err, (double)c1) * (1 - prior) / 2;
    double gl11 = pow(err, (double)c0) * pow((1 - err), (double)c1) * (1 - prior) / 2;
    double gl01 = pow(0.5, (double)(c0 + c1)) * prior;
    double lp[3] = {log10(gl00), log10(gl01), log10(gl11)};
    double m = lp[0];
    if (lp[1] > m) m = lp[1];
    if (lp[2] > m) m = lp[2];
    double s = 0.0;
    for (int i = 0; i < 3; i++) s = s + pow(10.0, lp[i] - m);
    double lse = m + log10(s);
    double prob[3], P[3];
    for (int i = 0; i < 3; i++) {
        prob[i] = lp[i] - lse;
        if (prob[i] > 0.0) prob[i] = 0.0; /* np.minimum(x, 0.0) */
        P[i] = pow(10.0, prob[i]);
    }
    for (int i = 0; i < 3; i++) g->pl[i] = (int)o_rint(-10 * log10(P[i]));
    int gq0 = (int)(-10 * log10(P[1] + P[2]));
    int gq1 = (int)(-10 * log10(P[0] + P[2]));
    int gq2 = (int)(-10 * log10(P[0] + P[1]));
    int gq = gq0;
    if (gq1 > gq) gq = gq1;
    if (gq2 > gq) gq = gq2;
    g->gq = gq;
    g->qual = fabs(o_rint((-10 * log10(P[0])) * 10.0) / 10.0); /* np.around(x, 1) */
    int best = 0; /* prob.index(max(prob)) -> first maximum */
    if (prob[1] > prob[best]) best = 1;
    if (prob[2] > prob[best]) best = 2;
    g->gt = best;
}

/* cal_CIPOS, cuteSV_genotype.py:58-60: int(1.96 * std / num ** 0.5); num ** 0.5 is libm pow */
static int cal_cipos(double std, int num) { return (int)(1.96 * std / pow((double)num, 0.5)); }

/* numpy's pairwise summation of a contiguous float64 array (what np.std's umr_sum runs) */
static double np_pairwise_sum(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* np.std(list_of_ints): mean = sum/n; x = arr - mean; sqrt(sum(x*x)/n) */
static double np_std_i32(const int32_t* v, int64_t n, double* scratch) {
    int64_t s = 0;
    for (int64_t i = 0; i < n; i++) s += v[i];
    double mean = (double)s / (double)n;
    for (int64_t i = 0; i < n; i++) {
        double x = (double)v[i] - mean;
        scratch[i] = x * x;
    }
    double ret = np_pairwise_sum(scratch, n);
    ret = ret / (double)n;
    return sqrt(ret);
}

/* ------------------------------------------------------------------------------------------ */
/* containers                                                                                 */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t chrom, a, b, rid, c;
    int64_t idx;
} Sig;

typedef struct {
    csv_cand* cands;
    csv_geno* genos;
    int64_t n_cand, cap_cand;
    int32_t* names;
    int64_t n_names, cap_names;
} Out;

static void out_init(Out* o) { memset(o, 0, sizeof(*o)); }
static void out_free(Out* o) {
    free(o->cands);
    free(o->genos);
    free(o->names);
}
static csv_cand* out_push(Out* o) {
    if (o->n_cand == o->cap_cand) {
        o->cap_cand = o->cap_cand ? o->cap_cand * 2 : 256;
        o->cands = (csv_cand*)realloc(o->cands, o->cap_cand * sizeof(csv_cand));
        o->genos = (csv_geno*)realloc(o->genos, o->cap_cand * sizeof(csv_geno));
    }
    csv_cand* c = &o->cands[o->n_cand];
    memset(c, 0, sizeof(*c));
    csv_geno* g = &o->genos[o->n_cand];
    memset(g, 0, sizeof(*g));
    g->dr = -1; g->gt = -1; g->status = 1;
    o->n_cand++;
    return c;
}
static int32_t out_names(Out* o, const int32_t* v, int64_t n) {
    if (o->n_names + n > o->cap_names) {
        while (o->n_names + n > o->cap_names) o->cap_names = o->cap_names ? o->cap_names * 2 : 1024;
        o->names = (int32_t*)realloc(o->names, o->cap_names * sizeof(int32_t));
    }
    memcpy(o->names + o->n_names, v, n * sizeof(int32_t));
    int32_t off = (int32_t)o->n_names;
    o->n_names += n;
    return off;
}

/* generic stable merge sort on an index/record array */
typedef int (*cmp_fn)(const void*, const void*);
static void msort(void* base, size_t n, size_t sz, cmp_fn cmp) {
    if (n < 2) return;
    char* tmp = (char*)malloc(n * sz);
    char* a = (char*)base;
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                if (cmp(a + j * sz, a + i * sz) < 0) memcpy(tmp + (k++) * sz, a + (j++) * sz, sz);
                else memcpy(tmp + (k++) * sz, a + (i++) * sz, sz);
            }
            while (i < mid) memcpy(tmp + (k++) * sz, a + (i++) * sz, sz);
            while (j < hi) memcpy(tmp + (k++) * sz, a + (j++) * sz, sz);
        }
        memcpy(a, tmp, n * sz);
    }
    free(tmp);
}

static int cmp_i32(const void* x, const void* y) {
    int32_t a = *(const int32_t*)x, b = *(const int32_t*)y;
    return (a > b) - (a < b);
}
/* len(set(ids)): number of distinct values (destroys order of tmp) */
static int64_t count_distinct(const int32_t* ids, int64_t n, int32_t* tmp) {
    if (n == 0) return 0;
    memcpy(tmp, ids, n * sizeof(int32_t));
    qsort(tmp, n, sizeof(int32_t), cmp_i32);
    int64_t u = 1;
    for (int64_t i = 1; i < n; i++)
        if (tmp[i] != tmp[i - 1]) tmp[u++] = tmp[i];
    return u; /* tmp[0..u) = sorted distinct ids */
}

/* ------------------------------------------------------------------------------------------ */
/* sort keys of process_process_sigs_type (cuteSV:764-801) + remove_duplicates_sorted (958-969) */
/* ------------------------------------------------------------------------------------------ */
#define CMP(x, y) do { if ((x) < (y)) return -1; if ((x) > (y)) return 1; } while (0)

/* DEL: (chr, int(pos), len, name) cuteSV:764; DUP: (chr, int(p1), int(p2), name) cuteSV:783 */
static int cmp_del(const void* p, const void* q) {
    const Sig *x = (const Sig*)p, *y = (const Sig*)q;
    CMP(x->chrom, y->chrom); CMP(x->a, y->a); CMP(x->b, y->b); CMP(x->rid, y->rid);
    return 0;
}
/* INS: (chr, int(pos), len, name, seq) cuteSV:774; a = 2*pos; seq itself is not available here:
 * ties keep input order (stable sort), see DESIGN.md "INS seq tie-break". */
static int cmp_ins(const void* p, const void* q) {
    const Sig *x = (const Sig*)p, *y = (const Sig*)q;
    CMP(x->chrom, y->chrom); CMP(x->a >> 1, y->a >> 1); CMP(x->b, y->b); CMP(x->rid, y->rid);
    return 0;
}
/* INV: (chr, strand, int(bp1), bp2, name) cuteSV:792 ("++" < "--" as strings: '+'=43 < '-'=45) */
static int cmp_inv(const void* p, const void* q) {
    const Sig *x = (const Sig*)p, *y = (const Sig*)q;
    CMP(x->chrom, y->chrom); CMP(x->c, y->c); CMP(x->a, y->a); CMP(x->b, y->b); CMP(x->rid, y->rid);
    return 0;
}
/* TRA: (chr1, chr2, type, int(pos1), pos2, name, "TRA") cuteSV:801; c = chr2*4+type */
static int cmp_tra(const void* p, const void* q) {
    const Sig *x = (const Sig*)p, *y = (const Sig*)q;
    CMP(x->chrom, y->chrom); CMP(x->c, y->c); CMP(x->a, y->a); CMP(x->b, y->b); CMP(x->rid, y->rid);
    return 0;
}
static int sig_equal(const Sig* x, const Sig* y) {
    return x->chrom == y->chrom && x->a == y->a && x->b == y->b && x->rid == y->rid && x->c == y->c;
}

static Sig* load_sorted(const csv_sig_cols* s, int svtype, int64_t* n_out) {
    int64_t n = s->n;
    Sig* v = (Sig*)malloc((n ? n : 1) * sizeof(Sig));
    for (int64_t i = 0; i < n; i++) {
        v[i].chrom = s->chrom[i]; v[i].a = s->a[i]; v[i].b = s->b[i]; v[i].rid = s->read_id[i];
        v[i].c = s->c ? s->c[i] : 0;
        v[i].idx = i;
    }
    cmp_fn f = svtype == CSV_INS ? cmp_ins : svtype == CSV_INV ? cmp_inv : svtype == CSV_TRA ? cmp_tra : cmp_del;
    msort(v, n, sizeof(Sig), f);
    /* remove_duplicates_sorted: adjacent full-tuple equality */
    int64_t j = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i == 0) { v[j] = v[i]; continue; }
        if (!sig_equal(&v[i], &v[j])) v[++j] = v[i];
    }
    *n_out = n ? j + 1 : 0;
    return v;
}

/* ------------------------------------------------------------------------------------------ */
/* reads table helpers for overlap_cover (cuteSV_genotype.py:95-159)                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t start, end, rid; } Prim;
typedef struct {
    int64_t n;            /* primaries on this contig */
    int64_t n_rows;       /* all reads-table rows on this contig */
    Prim* by_start;       /* sorted by start */
    int32_t* ends_sorted; /* ends, sorted */
    Prim* by_rid;         /* sorted by rid */
} ChromReads;

static int cmp_prim_start(const void* p, const void* q) { CMP(((const Prim*)p)->start, ((const Prim*)q)->start); return 0; }
static int cmp_prim_rid(const void* p, const void* q) { CMP(((const Prim*)p)->rid, ((const Prim*)q)->rid); return 0; }

static ChromReads* build_reads(const csv_reads_cols* r, int32_t n_contigs) {
    ChromReads* cr = (ChromReads*)calloc(n_contigs ? n_contigs : 1, sizeof(ChromReads));
    if (!r) return cr;
    for (int64_t i = 0; i < r->n; i++) {
        cr[r->chrom[i]].n_rows++;
        if (r->is_primary[i]) cr[r->chrom[i]].n++;
    }
    for (int32_t c = 0; c < n_contigs; c++) {
        cr[c].by_start = (Prim*)malloc((cr[c].n ? cr[c].n : 1) * sizeof(Prim));
        cr[c].n = 0;
    }
    for (int64_t i = 0; i < r->n; i++) {
        if (!r->is_primary[i]) continue;
        ChromReads* x = &cr[r->chrom[i]];
        x->by_start[x->n].start = r->start[i];
        x->by_start[x->n].end = r->end[i];
        x->by_start[x->n].rid = r->read_id[i];
        x->n++;
    }
    for (int32_t c = 0; c < n_contigs; c++) {
        ChromReads* x = &cr[c];
        int64_t n = x->n;
        x->by_rid = (Prim*)malloc((n ? n : 1) * sizeof(Prim));
        x->ends_sorted = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
        memcpy(x->by_rid, x->by_start, n * sizeof(Prim));
        qsort(x->by_start, n, sizeof(Prim), cmp_prim_start);
        qsort(x->by_rid, n, sizeof(Prim), cmp_prim_rid);
        for (int64_t i = 0; i < n; i++) x->ends_sorted[i] = x->by_start[i].end;
        qsort(x->ends_sorted, n, sizeof(int32_t), cmp_i32);
    }
    return cr;
}
static void free_reads(ChromReads* cr, int32_t n_contigs) {
    for (int32_t c = 0; c < n_contigs; c++) { free(cr[c].by_start); free(cr[c].ends_sorted); free(cr[c].by_rid); }
    free(cr);
}

/* Number of primaries r with r.start <= s and r.end >= e (integer window, e > s).
 * overlap_cover's event order (sv-right 0 < read-left 1 < read-right 2 < sv-left 3,
 * cuteSV_genotype.py:100-138) makes the cover set exactly this predicate; half-integer windows
 * (DUP/INV bias/2) are mapped by the caller to s=floor, e=ceil which is equivalent for integer
 * read coordinates.  Counted as A - B + C: A=#{start<=s}, B=#{end<e}, C=#{start>s and end<e}. */
static int64_t cover_count(const ChromReads* x, int64_t s, int64_t e) {
    int64_t n = x->n, lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) / 2; if (x->by_start[m].start <= s) lo = m + 1; else hi = m; }
    int64_t A = lo;
    lo = 0; hi = n;
    while (lo < hi) { int64_t m = (lo + hi) / 2; if (x->ends_sorted[m] < e) lo = m + 1; else hi = m; }
    int64_t B = lo, C = 0;
    for (int64_t i = A; i < n && x->by_start[i].start < e; i++)
        if (x->by_start[i].end < e) C++;
    return A - B + C;
}
static const Prim* find_prim(const ChromReads* x, int32_t rid) {
    int64_t lo = 0, hi = x->n;
    while (lo < hi) { int64_t m = (lo + hi) / 2; if (x->by_rid[m].rid < rid) lo = m + 1; else hi = m; }
    if (lo < x->n && x->by_rid[lo].rid == rid) return &x->by_rid[lo];
    return NULL;
}
static int covers(const Prim* p, int64_t s, int64_t e) { return p && p->start <= s && p->end >= e; }

/* assign_gt (cuteSV_genotype.py:161-173) for one window */
static void genotype_one(const ChromReads* x, int64_t s, int64_t e, const int32_t* names, int64_t n_names, csv_geno* g) {
    int64_t cover = cover_count(x, s, e), sup = 0;
    for (int64_t i = 0; i < n_names; i++)
        if (covers(find_prim(x, names[i]), s, e)) sup++;
    csvo_cal_gl((int)(cover - sup), (int)n_names, g);
}
/* DUP / INV: union of the covers of two windows (resolveDUP.py:146-159, resolveINV.py:218-229) */
static void genotype_two(const ChromReads* x, int64_t s1, int64_t e1, int64_t s2, int64_t e2,
                         const int32_t* names, int64_t n_names, csv_geno* g) {
    int64_t smin = s1 < s2 ? s1 : s2, emax = e1 > e2 ? e1 : e2;
    int64_t cover = cover_count(x, s1, e1) + cover_count(x, s2, e2) - cover_count(x, smin, emax);
    int64_t sup = 0;
    for (int64_t i = 0; i < n_names; i++) {
        const Prim* p = find_prim(x, names[i]);
        if (covers(p, s1, e1) || covers(p, s2, e2)) sup++;
    }
    csvo_cal_gl((int)(cover - sup), (int)n_names, g);
}
static int64_t floor_half(int64_t twice) { return twice >= 0 ? twice / 2 : -((-twice + 1) / 2); }
static int64_t ceil_half(int64_t twice) { return twice >= 0 ? (twice + 1) / 2 : -((-twice) / 2); }

/* ------------------------------------------------------------------------------------------ */
/* INS / DEL: generate_del_cluster / generate_ins_cluster (resolveINDEL.py:110-219, 319-432)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t pos, len, rid, seqlen; int64_t idx; int64_t first; } Mem;
typedef struct { int32_t rid; int64_t order; } RidOrd;
static int cmp_ridord(const void* p, const void* q) {
    const RidOrd *x = (const RidOrd*)p, *y = (const RidOrd*)q;
    CMP(x->rid, y->rid); CMP(x->order, y->order); return 0;
}
static int cmp_mem_first(const void* p, const void* q) { CMP(((const Mem*)p)->first, ((const Mem*)q)->first); return 0; }
static int cmp_mem_len(const void* p, const void* q) { CMP(((const Mem*)p)->len, ((const Mem*)q)->len); return 0; }
typedef struct { int64_t start, count, ord; } Allele;
static int cmp_allele(const void* p, const void* q) { CMP(((const Allele*)p)->count, ((const Allele*)q)->count); return 0; }
typedef struct { double d; int64_t i; } VarEnt;
static int cmp_var(const void* p, const void* q) { CMP(((const VarEnt*)p)->d, ((const VarEnt*)q)->d); return 0; }

static void generate_indel_cluster(const Sig* cl, int64_t m, int svtype, const csv_params* P, int32_t cluster_ord, Out* out) {
    double ratio = svtype == CSV_INS ? P->ratio_ins : P->ratio_del;
    double keep = P->remain_reads_ratio > 1 ? 1 : P->remain_reads_ratio; /* resolveINDEL.py:46-47 */
    /* Remove duplicates per read: dict keeps first-occurrence order, value replaced when the
       new element is strictly longer (resolveINDEL.py:125-131) */
    RidOrd* ro = (RidOrd*)malloc(m * sizeof(RidOrd));
    for (int64_t i = 0; i < m; i++) { ro[i].rid = cl[i].rid; ro[i].order = i; }
    qsort(ro, m, sizeof(RidOrd), cmp_ridord);
    Mem* mem = (Mem*)malloc(m * sizeof(Mem));
    int64_t u = 0;
    for (int64_t i = 0; i < m;) {
        int64_t j = i, best = ro[i].order;
        while (j < m && ro[j].rid == ro[i].rid) {
            if (cl[ro[j].order].b > cl[best].b) best = ro[j].order;
            j++;
        }
        const Sig* s = &cl[best];
        mem[u].pos = svtype == CSV_INS ? (s->a >> 1) : s->a;
        mem[u].len = s->b; mem[u].rid = s->rid; mem[u].seqlen = s->c; mem[u].idx = s->idx;
        mem[u].first = ro[i].order;
        u++;
        i = j;
    }
    free(ro);
    if (u < P->min_support) { free(mem); return; } /* :133 */
    msort(mem, u, sizeof(Mem), cmp_mem_first);      /* dict order */
    msort(mem, u, sizeof(Mem), cmp_mem_len);        /* sorted(..., key=len) :136 (stable) */
    int64_t sum_len = 0;
    for (int64_t i = 0; i < u; i++) sum_len += mem[i].len;
    double thr = ratio * ((double)sum_len / (double)u); /* :138 */
    Allele* al = (Allele*)malloc(u * sizeof(Allele));
    int64_t na = 0;
    al[0].start = 0; al[0].count = 1; al[0].ord = 0; na = 1;
    int32_t last_len = mem[0].len;
    for (int64_t i = 1; i < u; i++) {
        if ((double)(mem[i].len - last_len) > thr) { al[na].start = i; al[na].count = 0; al[na].ord = na; na++; }
        al[na - 1].count++;
        last_len = mem[i].len;
    }
    msort(al, na, sizeof(Allele), cmp_allele); /* sorted(allele_collect, key=[support]) :163 */
    int32_t* ipos = (int32_t*)malloc(u * sizeof(int32_t));
    int32_t* ilen = (int32_t*)malloc(u * sizeof(int32_t));
    int32_t* irid = (int32_t*)malloc(u * sizeof(int32_t));
    double* scratch = (double*)malloc(u * sizeof(double));
    VarEnt* var = (VarEnt*)malloc(u * sizeof(VarEnt));
    for (int64_t k = 0; k < na; k++) {
        int64_t n = al[k].count;
        if (n < P->min_support_allele) continue; /* :166 */
        const Mem* a = mem + al[k].start;
        for (int64_t i = 0; i < n; i++) { ipos[i] = a[i].pos; ilen[i] = a[i].len; irid[i] = a[i].rid; }
        int64_t remain = (int64_t)(keep * (double)n);
        if (remain < 1) remain = 1;
        int64_t sp = 0, sl = 0;
        for (int64_t i = 0; i < n; i++) { sp += ipos[i]; sl += ilen[i]; }
        double pos_mean = (double)sp / (double)n;
        for (int64_t i = 0; i < n; i++) { var[i].d = fabs((double)ipos[i] - pos_mean); var[i].i = i; }
        msort(var, n, sizeof(VarEnt), cmp_var);
        int64_t s = 0;
        for (int64_t i = 0; i < remain; i++) s += ipos[var[i].i];
        double breakpointStart = (double)s / (double)remain;
        int32_t search_threshold = ipos[var[0].i];
        double len_mean = (double)sl / (double)n;
        for (int64_t i = 0; i < n; i++) { var[i].d = fabs((double)ilen[i] - len_mean); var[i].i = i; }
        msort(var, n, sizeof(VarEnt), cmp_var);
        s = 0;
        for (int64_t i = 0; i < remain; i++) s += ilen[var[i].i];
        double signalLen = (double)s / (double)remain;
        int cipos = cal_cipos(np_std_i32(ipos, n, scratch), (int)n);
        int cilen = cal_cipos(np_std_i32(ilen, n, scratch), (int)n);
        int32_t aux = 0;
        int32_t pos_out = (int32_t)breakpointStart;
        if (svtype == CSV_INS) {
            int found = 0;
            for (int64_t i = 0; i < n; i++) { /* resolveINDEL.py:399-405 */
                if (a[i].seqlen >= (int32_t)signalLen) { pos_out = a[i].pos; aux = (int32_t)a[i].idx; found = 1; break; }
            }
            if (!found) continue;
            search_threshold = pos_out;
        }
        csv_cand* c = out_push(out);
        c->svtype = svtype; c->chrom = cl[0].chrom; c->pos = pos_out;
        c->len = svtype == CSV_DEL ? (int32_t)(-signalLen) : (int32_t)signalLen;
        c->support = (int32_t)n; c->cipos = cipos; c->cilen = cilen; c->search_pos = search_threshold;
        c->aux = aux; c->cluster = cluster_ord;
        c->names_cnt = (int32_t)n;
        c->names_off = out_names(out, irid, n);
    }
    free(ipos); free(ilen); free(irid); free(scratch); free(var); free(al); free(mem);
}

/* resolution_DEL / resolution_INS sweep (resolveINDEL.py:55-100, 261-310) + call_gt (441-479) */
static void resolve_indel(const Sig* v, int64_t n, int svtype, const csv_params* P, const ChromReads* cr, Out* out) {
    int32_t bias = svtype == CSV_INS ? P->bias_ins : P->bias_del;
    int64_t first_cand = out->n_cand;
    int32_t ord = 0;
    int64_t lo = 0;
    for (int64_t i = 1; i <= n; i++) {
        int boundary = i == n;
        if (!boundary) {
            int32_t p = svtype == CSV_INS ? (v[i].a >> 1) : v[i].a;
            int32_t q = svtype == CSV_INS ? (v[i - 1].a >> 1) : v[i - 1].a;
            boundary = p - q > bias;
        }
        if (boundary) {
            if (i - lo >= P->min_support) generate_indel_cluster(v + lo, i - lo, svtype, P, ord, out);
            ord++;
            lo = i;
        }
    }
    if (P->genotype) {
        int32_t gb = svtype == CSV_INS ? P->gt_bias_ins : P->bias_del;
        for (int64_t k = first_cand; k < out->n_cand; k++) {
            csv_cand* c = &out->cands[k];
            if (cr->n_rows == 0) { c->flags |= CSV_F_NO_READS; continue; }
            int64_t s = (int64_t)c->search_pos - gb; if (s < 0) s = 0;
            int64_t e = (int64_t)c->search_pos + gb;
            genotype_one(cr, s, e, out->names + c->names_off, c->names_cnt, &out->genos[k]);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* DUP: resolveDUP.py:17-181                                                                   */
/* ------------------------------------------------------------------------------------------ */
static int cmp_sig_b(const void* p, const void* q) { CMP(((const Sig*)p)->b, ((const Sig*)q)->b); return 0; }

static void generate_dup_cluster(const Sig* cl_in, int64_t m, const csv_params* P, int32_t ord, Out* out) {
    int32_t* tmp = (int32_t*)malloc(m * sizeof(int32_t));
    int32_t* ids = (int32_t*)malloc(m * sizeof(int32_t));
    for (int64_t i = 0; i < m; i++) ids[i] = cl_in[i].rid;
    if (count_distinct(ids, m, tmp) < P->min_support) { free(tmp); free(ids); return; } /* :82-84 */
    Sig* cl = (Sig*)malloc(m * sizeof(Sig));
    memcpy(cl, cl_in, m * sizeof(Sig));
    msort(cl, m, sizeof(Sig), cmp_sig_b); /* :86 stable sort by pos2 */
    for (int64_t lo = 0; lo < m;) {
        int64_t hi = lo + 1;
        while (hi < m && !(cl[hi].b - cl[hi - 1].b > P->bias_dup)) hi++; /* :90-94 */
        int64_t n = hi - lo;
        for (int64_t i = 0; i < n; i++) ids[i] = cl[lo + i].rid;
        int64_t u = count_distinct(ids, n, tmp);
        if (u >= P->min_support) {
            int64_t low_b = (int64_t)((double)n * 0.4), up_b = (int64_t)((double)n * 0.6);
            int64_t bp1, bp2;
            if (low_b == up_b) { bp1 = cl[lo + low_b].a; bp2 = cl[lo + low_b].b; }
            else {
                int64_t s1 = 0, s2 = 0;
                for (int64_t i = low_b; i < up_b; i++) { s1 += cl[lo + i].a; s2 += cl[lo + i].b; }
                bp1 = (int64_t)((double)s1 / (double)(up_b - low_b));
                bp2 = (int64_t)((double)s2 / (double)(up_b - low_b));
            }
            int64_t d = bp2 - bp1;
            if ((P->min_size <= d && d <= P->max_size) || (P->min_size <= d && P->max_size == -1)) { /* :112 */
                csv_cand* c = out_push(out);
                c->svtype = CSV_DUP; c->chrom = cl[0].chrom; c->pos = (int32_t)bp1; c->pos2 = (int32_t)bp2;
                c->len = (int32_t)d; c->support = (int32_t)u; c->cluster = ord;
                c->names_cnt = (int32_t)u;
                c->names_off = out_names(out, tmp, u); /* list(set(...)): order unspecified -> sorted ids */
            }
        }
        lo = hi;
    }
    free(cl); free(tmp); free(ids);
}

static void resolve_dup(const Sig* v, int64_t n, const csv_params* P, const ChromReads* cr, Out* out) {
    int64_t first_cand = out->n_cand, lo = 0;
    int32_t ord = 0;
    for (int64_t i = 1; i <= n; i++) {
        int boundary = i == n || (v[i].a - v[i - 1].a > P->bias_dup); /* :35 */
        if (boundary) {
            if (i - lo >= P->min_support) generate_dup_cluster(v + lo, i - lo, P, ord, out);
            ord++;
            lo = i;
        }
    }
    if (P->genotype) { /* call_gt resolveDUP.py:137-181 */
        for (int64_t k = first_cand; k < out->n_cand; k++) {
            csv_cand* c = &out->cands[k];
            if (cr->n_rows == 0) { c->flags |= CSV_F_NO_READS; continue; }
            int64_t nb = c->pos2 - c->pos; if (P->bias_dup < nb) nb = P->bias_dup; /* min(bias, len) */
            int64_t s1 = floor_half(2 * (int64_t)c->pos - nb); if (2 * (int64_t)c->pos - nb < 0) s1 = 0;
            int64_t e1 = ceil_half(2 * (int64_t)c->pos + nb);
            int64_t s2 = floor_half(2 * (int64_t)c->pos2 - nb); if (2 * (int64_t)c->pos2 - nb < 0) s2 = 0;
            int64_t e2 = ceil_half(2 * (int64_t)c->pos2 + nb);
            genotype_two(cr, s1, e1, s2, e2, out->names + c->names_off, c->names_cnt, &out->genos[k]);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* INV: resolveINV.py:6-252                                                                    */
/* ------------------------------------------------------------------------------------------ */
static void inv_emit(const Sig* sub, int64_t cnt, const csv_params* P, int32_t ord, Out* out,
                     int64_t sum1, int64_t sum2, int32_t* ids, int32_t* tmp) {
    if (cnt < P->min_support) return; /* temp_count >= read_count :126,173 */
    /* temp_id: distinct names in first-occurrence order */
    int64_t u = 0;
    for (int64_t i = 0; i < cnt; i++) {
        int seen = 0;
        for (int64_t j = 0; j < u; j++) if (ids[j] == sub[i].rid) { seen = 1; break; }
        if (!seen) ids[u++] = sub[i].rid;
    }
    (void)tmp;
    int64_t bp1 = (int64_t)o_rint((double)sum1 / (double)cnt); /* round(): half-to-even :129 */
    int64_t bp2 = (int64_t)o_rint((double)sum2 / (double)cnt);
    int64_t inv_len = bp2 - bp1;
    if (inv_len >= P->min_size && u >= P->min_support) {
        if (inv_len <= P->max_size || P->max_size == -1) {
            csv_cand* c = out_push(out);
            c->svtype = CSV_INV; c->chrom = sub[0].chrom; c->pos = (int32_t)bp1; c->pos2 = (int32_t)bp2;
            c->len = (int32_t)inv_len; c->support = (int32_t)u; c->aux = sub[0].c; c->cluster = ord;
            c->names_cnt = (int32_t)u;
            c->names_off = out_names(out, ids, u);
        }
    }
}

static void generate_inv_cluster(const Sig* cl_in, int64_t m, const csv_params* P, int32_t ord, Out* out) {
    int32_t* tmp = (int32_t*)malloc(m * sizeof(int32_t));
    int32_t* ids = (int32_t*)malloc(m * sizeof(int32_t));
    for (int64_t i = 0; i < m; i++) ids[i] = cl_in[i].rid;
    if (count_distinct(ids, m, tmp) < P->min_support) { free(tmp); free(ids); return; } /* :106-109 */
    Sig* cl = (Sig*)malloc(m * sizeof(Sig));
    memcpy(cl, cl_in, m * sizeof(Sig));
    msort(cl, m, sizeof(Sig), cmp_sig_b); /* :111 */
    for (int64_t lo = 0; lo < m;) {
        int64_t hi = lo + 1, s1 = cl[lo].a, s2 = cl[lo].b;
        while (hi < m && !(cl[hi].b - cl[hi - 1].b > P->bias_inv)) { s1 += cl[hi].a; s2 += cl[hi].b; hi++; }
        inv_emit(cl + lo, hi - lo, P, ord, out, s1, s2, ids, tmp);
        lo = hi;
    }
    free(cl); free(tmp); free(ids);
}

static void resolve_inv(const Sig* v, int64_t n, const csv_params* P, const ChromReads* cr, Out* out) {
    int64_t first_cand = out->n_cand, lo = 0;
    int32_t ord = 0;
    for (int64_t i = 1; i <= n; i++) {
        int boundary = i == n || (v[i].a - v[i - 1].a > P->bias_inv) || (v[i].b - v[i - 1].b > P->bias_inv) ||
                       v[i].c != v[i - 1].c; /* :56 */
        if (boundary) {
            if (i - lo >= P->min_support) generate_inv_cluster(v + lo, i - lo, P, ord, out);
            ord++;
            lo = i;
        }
    }
    if (P->genotype) { /* call_gt resolveINV.py:208-252 */
        for (int64_t k = first_cand; k < out->n_cand; k++) {
            csv_cand* c = &out->cands[k];
            if (cr->n_rows == 0) { c->flags |= CSV_F_NO_READS; continue; }
            int64_t nb = P->bias_inv;
            int64_t s1 = floor_half(2 * (int64_t)c->pos - nb); if (2 * (int64_t)c->pos - nb < 0) s1 = 0;
            int64_t e1 = ceil_half(2 * (int64_t)c->pos + nb);
            int64_t s2 = floor_half(2 * (int64_t)c->pos2 - nb); if (2 * (int64_t)c->pos2 - nb < 0) s2 = 0;
            int64_t e2 = ceil_half(2 * (int64_t)c->pos2 + nb);
            genotype_two(cr, s1, e1, s2, e2, out->names + c->names_off, c->names_cnt, &out->genos[k]);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* TRA: resolveTRA.py:30-255 (call_gt needs the BAM and stays on the host: CSV_F_GT_HOST)      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int64_t s1, s2, listlen, distinct, lo, hi, ord; } TraSub;
static int cmp_trasub(const void* p, const void* q) { CMP(-((const TraSub*)p)->distinct, -((const TraSub*)q)->distinct); return 0; }

static void tra_emit(const Sig* cl, const TraSub* t, int32_t chr1, int32_t c_aux, const csv_params* P, int32_t ord,
                     Out* out, int32_t* ids, int32_t* tmp) {
    int64_t n = t->hi - t->lo;
    for (int64_t i = 0; i < n; i++) ids[i] = cl[t->lo + i].rid;
    int64_t u = count_distinct(ids, n, tmp);
    csv_cand* c = out_push(out);
    c->svtype = CSV_TRA; c->chrom = chr1;
    c->pos = (int32_t)((double)t->s1 / (double)t->listlen);  /* int(temp[k][0]/len(temp[k][2])) :173 */
    c->pos2 = (int32_t)((double)t->s2 / (double)t->listlen);
    c->support = (int32_t)u; c->aux = c_aux; c->cluster = ord;
    c->flags = P->genotype ? CSV_F_GT_HOST : 0;
    c->names_cnt = (int32_t)u;
    c->names_off = out_names(out, tmp, u); /* set(): order unspecified -> sorted ids */
}

static void generate_tra_cluster(const Sig* cl_in, int64_t m, const csv_params* P, int32_t ord, Out* out) {
    Sig* cl = (Sig*)malloc(m * sizeof(Sig));
    memcpy(cl, cl_in, m * sizeof(Sig));
    msort(cl, m, sizeof(Sig), cmp_sig_b); /* :109 sorted by pos2 */
    int32_t* tmp = (int32_t*)malloc(m * sizeof(int32_t));
    int32_t* ids = (int32_t*)malloc(m * sizeof(int32_t));
    TraSub* sub = (TraSub*)malloc(m * sizeof(TraSub));
    int64_t ns = 0;
    /* :113-124: temp starts with element 0, then the loop visits element 0 AGAIN */
    sub[0].s1 = cl[0].a; sub[0].s2 = cl[0].b; sub[0].listlen = 1; sub[0].lo = 0; sub[0].hi = 0; sub[0].ord = 0; ns = 1;
    int32_t last = cl[0].b;
    for (int64_t i = 0; i < m; i++) {
        if (cl[i].b - last > P->bias_tra) {
            sub[ns - 1].hi = i;
            sub[ns].s1 = cl[i].a; sub[ns].s2 = cl[i].b; sub[ns].listlen = 1; sub[ns].lo = i; sub[ns].ord = ns; ns++;
        } else {
            sub[ns - 1].s1 += cl[i].a; sub[ns - 1].s2 += cl[i].b; sub[ns - 1].listlen++;
        }
        last = cl[i].b;
    }
    sub[ns - 1].hi = m;
    for (int64_t i = 0; i < m; i++) ids[i] = cl[i].rid;
    if (count_distinct(ids, m, tmp) < P->min_support) goto done; /* :128 */
    for (int64_t k = 0; k < ns; k++) {
        int64_t n = sub[k].hi - sub[k].lo;
        for (int64_t i = 0; i < n; i++) ids[i] = cl[sub[k].lo + i].rid;
        sub[k].distinct = count_distinct(ids, n, tmp);
    }
    msort(sub, ns, sizeof(TraSub), cmp_trasub); /* :131 */
    {
        int32_t chr1 = cl[0].chrom, aux = cl_in[0].c; /* BND_type = semi_tra_cluster[0][3] :108 */
        if (ns > 1 && (double)sub[1].distinct >= 0.5 * (double)P->min_support) { /* :133 */
            if ((double)(sub[0].distinct + sub[1].distinct) >= (double)m * P->ratio_tra) { /* :134 */
                tra_emit(cl, &sub[0], chr1, aux, P, ord, out, ids, tmp);
                tra_emit(cl, &sub[1], chr1, aux, P, ord, out, ids, tmp);
            }
        } else {
            if ((double)sub[0].distinct >= (double)m * P->ratio_tra) /* :211 */
                tra_emit(cl, &sub[0], chr1, aux, P, ord, out, ids, tmp);
        }
    }
done:
    free(cl); free(tmp); free(ids); free(sub);
}

static void resolve_tra(const Sig* v, int64_t n, const csv_params* P, Out* out) {
    int64_t lo = 0;
    int32_t ord = 0;
    for (int64_t i = 1; i <= n; i++) {
        /* :41 chr2 change flushes; :65 pos1 gap or BND type change */
        int boundary = i == n || v[i].c != v[i - 1].c || (v[i].a - v[i - 1].a > P->bias_tra);
        if (boundary) {
            if (i - lo >= P->min_support) generate_tra_cluster(v + lo, i - lo, P, ord, out);
            ord++;
            lo = i;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* TRA genotyping: call_gt (resolveTRA.py:260-309), count_coverage / threshold_ref_count        */
/* (cuteSV_genotype.py:62-93) over the packed all-alignments table (BAM order)                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int64_t n; const int32_t *chrom, *start, *end, *rid; const uint8_t* prim; int64_t* off; } AlnTab;

static int threshold_ref_count(int num) {
    if (num <= 2) return 20 * num;
    if (num <= 5) return 9 * num;
    if (num <= 15) return 7 * num;
    return 5 * num;
}
static int in_sorted(const int32_t* v, int n, int32_t x) {
    int lo = 0, hi = n;
    while (lo < hi) { int m = (lo + hi) / 2; if (v[m] < x) lo = m + 1; else hi = m; }
    return lo < n && v[lo] == x;
}
/* querydata is a set of names; with one primary record per name its size is a plain count */
/* (xs, xe): window whose spanning reads are already in the set (same-contig second region); xs > xe: none */
static int count_coverage(const AlnTab* A, int chr, int64_t s, int64_t e, const int32_t* sup, int n_sup, int up_bound, int itround,
                          int* nset, int* dr, int64_t xs, int64_t xe) {
    int64_t iteration = 0, primary = 0;
    for (int64_t i = A->off[chr]; i < A->off[chr + 1]; i++) {
        if (A->start[i] >= e) break;             /* f.fetch(chr, s, e): start < e and end > s, in BAM order */
        if (!(A->end[i] > s)) continue;
        iteration += 1;
        if (!A->prim[i]) continue;               /* i.flag not in [0, 16] */
        primary += 1;
        if (A->start[i] < s && A->end[i] > e) {
            int seen = xs <= xe && A->start[i] < xs && A->end[i] > xe; /* read_count.add() of a name already in the set */
            if (!seen) {
                *nset += 1;
                if (!in_sorted(sup, n_sup, A->rid[i])) *dr += 1;
            }
            if (*nset >= up_bound) return 1;
        }
        if (iteration >= itround) return ((double)primary / (double)iteration) <= 0.2 ? 1 : -1;
    }
    return 0;
}
static void tra_genotype(const AlnTab* A, const int64_t* contig_len, const csv_params* P, csv_cand* c, const int32_t* names, csv_geno* g) {
    int chr1 = c->chrom, chr2 = c->aux >> 2, n_sup = c->names_cnt;
    int up = threshold_ref_count(n_sup), nset = 0, dr = 0;
    int64_t s = (int64_t)c->pos - P->bias_tra; if (s < 0) s = 0;
    int64_t e = (int64_t)c->pos + P->bias_tra; if (e > contig_len[chr1]) e = contig_len[chr1];
    int st = count_coverage(A, chr1, s, e, names, n_sup, up, P->gt_round, &nset, &dr, 1, 0);
    int64_t s1 = s, e1 = e;
    c->flags &= ~CSV_F_GT_HOST;
    if (st == -1) { g->dr = -1; g->dv = n_sup; g->gt = -1; g->pl[0] = g->pl[1] = g->pl[2] = 0; g->gq = 0; g->status = 2; g->qual = 0.0; return; }
    if (st == 0) {
        s = (int64_t)c->pos2 - P->bias_tra; if (s < 0) s = 0;
        e = (int64_t)c->pos2 + P->bias_tra; if (e > contig_len[chr2]) e = contig_len[chr2];
        if (chr2 == chr1) count_coverage(A, chr2, s, e, names, n_sup, up, P->gt_round, &nset, &dr, s1, e1);
        else count_coverage(A, chr2, s, e, names, n_sup, up, P->gt_round, &nset, &dr, 1, 0);
    }
    csvo_cal_gl(dr, n_sup, g);
}

/* ------------------------------------------------------------------------------------------ */
/* public entry points                                                                         */
/* ------------------------------------------------------------------------------------------ */

/* Whole clustering phase (cuteSV:1113-1199) preceded by the rebuild sort (cuteSV:750-857).
 * Output order: svtype (DEL, INS, INV, DUP, TRA), contig id, reference emission order.
 * Returns 0, or CSV_E_CAPACITY with the needed sizes in n_cand/n_names. */
int csvo_cluster(const csv_params* P, int32_t n_contigs, const int64_t* contig_len, const csv_sig_cols sigs[CSV_NTYPES],
                 const csv_reads_cols* reads, uint32_t type_mask, csv_cand* cands, csv_geno* genos, int64_t cap_cand,
                 int32_t* names, int64_t cap_names, int64_t* n_cand, int64_t* n_names, int n_threads, const csv_reads_cols* aln) {
    ChromReads* cr = build_reads(P->genotype ? reads : NULL, n_contigs);
    Sig* sorted[CSV_NTYPES];
    int64_t ns[CSV_NTYPES];
    int64_t* chr_off[CSV_NTYPES];
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < CSV_NTYPES; t++) {
        sorted[t] = NULL; ns[t] = 0; chr_off[t] = NULL;
        if (!(type_mask >> t & 1)) continue;
        sorted[t] = load_sorted(&sigs[t], t, &ns[t]);
        chr_off[t] = (int64_t*)calloc(n_contigs + 1, sizeof(int64_t));
        for (int64_t i = 0; i < ns[t]; i++) chr_off[t][sorted[t][i].chrom + 1]++;
        for (int32_t c = 0; c < n_contigs; c++) chr_off[t][c + 1] += chr_off[t][c];
    }
    int64_t n_tasks = (int64_t)CSV_NTYPES * n_contigs;
    Out* outs = (Out*)malloc((n_tasks ? n_tasks : 1) * sizeof(Out));
    for (int64_t k = 0; k < n_tasks; k++) out_init(&outs[k]);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = 0; k < n_tasks; k++) {
        int t = (int)(k / n_contigs);
        int32_t c = (int32_t)(k % n_contigs);
        if (!sorted[t]) continue;
        const Sig* v = sorted[t] + chr_off[t][c];
        int64_t n = chr_off[t][c + 1] - chr_off[t][c];
        if (n == 0) continue;
        switch (t) {
            case CSV_DEL: case CSV_INS: resolve_indel(v, n, t, P, &cr[c], &outs[k]); break;
            case CSV_INV: resolve_inv(v, n, P, &cr[c], &outs[k]); break;
            case CSV_DUP: resolve_dup(v, n, P, &cr[c], &outs[k]); break;
            case CSV_TRA: resolve_tra(v, n, P, &outs[k]); break;
        }
    }
    int64_t tc = 0, tn = 0;
    for (int64_t k = 0; k < n_tasks; k++) { tc += outs[k].n_cand; tn += outs[k].n_names; }
    *n_cand = tc; *n_names = tn;
    int rc = CSV_OK;
    if (tc > cap_cand || tn > cap_names) rc = CSV_E_CAPACITY;
    else {
        int64_t oc = 0, on = 0;
        for (int64_t k = 0; k < n_tasks; k++) {
            for (int64_t i = 0; i < outs[k].n_cand; i++) {
                cands[oc] = outs[k].cands[i];
                cands[oc].names_off += (int32_t)on;
                genos[oc] = outs[k].genos[i];
                oc++;
            }
            if (outs[k].n_names) memcpy(names + on, outs[k].names, outs[k].n_names * sizeof(int32_t));
            on += outs[k].n_names;
        }
        if (P->genotype && aln && aln->n > 0) {  /* TRA rows: genotype from the all-alignments table */
            AlnTab A;
            A.n = aln->n; A.chrom = aln->chrom; A.start = aln->start; A.end = aln->end; A.rid = aln->read_id; A.prim = aln->is_primary;
            A.off = (int64_t*)calloc(n_contigs + 2, sizeof(int64_t));
            for (int64_t i = 0; i < aln->n; i++) A.off[aln->chrom[i] + 1]++;
            for (int32_t c = 0; c < n_contigs; c++) A.off[c + 1] += A.off[c];
            for (int64_t i = 0; i < oc; i++)
                if (cands[i].svtype == CSV_TRA) tra_genotype(&A, contig_len, P, &cands[i], names + cands[i].names_off, &genos[i]);
            free(A.off);
        }
    }
    for (int64_t k = 0; k < n_tasks; k++) out_free(&outs[k]);
    free(outs);
    for (int t = 0; t < CSV_NTYPES; t++) { free(sorted[t]); free(chr_off[t]); }
    free_reads(cr, n_contigs);
    return rc;
}

/* n ** 0.5 as Python evaluates it (libm pow), for the CIPOS table pin */
double csvo_pow_half(int n) { return pow((double)n, 0.5); }
double csvo_np_std_i32(const int32_t* v, int64_t n) {
    double* scratch = (double*)malloc((n ? n : 1) * sizeof(double));
    double r = np_std_i32(v, n, scratch);
    free(scratch);
    return r;
}
int csvo_cal_cipos(double std, int num) { return cal_cipos(std, num); }
