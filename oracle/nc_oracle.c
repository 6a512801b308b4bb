/* TEST INFRASTRUCTURE ONLY -- see nc_oracle.h for the rules and the parity status (PINNED
 * against the compiled reference headers by tests/test_oracle_vs_ref.py).
 *
 * Plain-C restatement of nanocall's decoding hot path.  Each function names the reference
 * file:line it follows (paths relative to /root/reference/src).  The code is written for
 * obviousness, not speed: dense N x 4096 matrices, explicit neighbour lists, sequential
 * accumulation in the reference's order.
 */
#define _GNU_SOURCE
#include "nc_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define S NCO_S
#define K NCO_K

/* ------------------------------------------------------------------ Kmer.hpp */

static unsigned prefix(unsigned i, unsigned k) { return i >> (2 * (K - k)); }          /* Kmer.hpp:69-72 */
static unsigned suffix(unsigned i, unsigned k) { return i & ((1u << (2 * k)) - 1); }   /* Kmer.hpp:73-76 */

unsigned nco_min_skip(unsigned k1, unsigned k2) /* Kmer.hpp:51-68 */
{
    if (k1 == k2) return 0;
    for (unsigned k = K - 1; k > 0; --k)
        if (suffix(k1, k) == prefix(k2, k)) return K - k;
    return K;
}

unsigned nco_max_self_overlap(unsigned i) /* Kmer.hpp:81-110 */
{
    for (unsigned k = K - 1; k >= 1; --k)
        if (suffix(i, k) == prefix(i, k)) return k;
    return 0;
}

void nco_kmer_string(unsigned k, char out[6]) /* Kmer.hpp:41-50 */
{
    static const char b[4] = { 'A', 'C', 'G', 'T' };
    for (unsigned j = 0; j < K; ++j) out[j] = b[(k >> (2 * (K - j - 1))) & 3];
}

/* ------------------------------------------------------------------ logsum.hpp */

static float g_tbl[NCO_LOGSUM_TBL];
static pthread_once_t g_tbl_once = PTHREAD_ONCE_INIT;
static void tbl_init(void) /* logsum.hpp:113-127: double log/exp, stored as float */
{
    for (int i = 0; i < NCO_LOGSUM_TBL; ++i) g_tbl[i] = (float)log(1. + exp((double)-i / 1000.f));
}
const float* nco_flogsum_table(void)
{
    pthread_once(&g_tbl_once, tbl_init);
    return g_tbl;
}
float nco_flogsum(float a, float b) /* logsum.hpp:141-154 */
{
    const float* tbl = nco_flogsum_table();
    const float max = a > b ? a : b;
    const float min = a < b ? a : b;
    return (min == -INFINITY || max - min >= (float)(NCO_LOGSUM_TBL - 1) / 1000.f)
        ? max
        : max + tbl[(int)((max - min) * 1000.f)];
}

/* ------------------------------------------------------------------ alg.hpp */

void nco_mean_stdv(unsigned n, const float* x, float* mean_out, float* stdv_out) /* alg.hpp:466-482 */
{
    float s = 0.0f, s2 = 0.0f;
    unsigned long cnt = 0;
    for (unsigned i = 0; i < n; ++i) { s += x[i]; s2 += x[i] * x[i]; ++cnt; }
    float mean = cnt > 0 ? s / cnt : 0.0f;
    /* (s2 - s*mean*2.0 + mean*mean*(float)n) / (n-1): the 2.0 promotes to double from there on */
    float stdv = cnt > 1 ? (float)sqrt(((double)s2 - (double)(s * mean) * 2.0 + (double)(mean * mean * (float)cnt)) / (double)(cnt - 1)) : 0.0f;
    *mean_out = mean;
    *stdv_out = stdv;
}

/* ------------------------------------------------------------------ Pore_Model.hpp */

static void model_stats(nco_model* m) /* Pore_Model.hpp:308-314 */
{
    nco_mean_stdv(S, m->level_mean, &m->mean, &m->stdv);
}

void nco_model_prepare(const float* table, nco_model* m) /* load_from_vector, Pore_Model.hpp:220-239 */
{
    for (unsigned i = 0; i < S; ++i)
    {
        m->level_mean[i] = table[4 * i + 0];
        m->level_stdv[i] = table[4 * i + 1];
        m->sd_mean[i] = table[4 * i + 2];
        m->sd_stdv[i] = table[4 * i + 3];
        /* update_sd_lambda (:112): pow in double, result narrowed */
        m->sd_lambda[i] = (float)(pow((double)m->sd_mean[i], 3.0) / pow((double)m->sd_stdv[i], 2.0));
        /* update_logs (:118-124): float log */
        m->log_level_stdv[i] = logf(m->level_stdv[i]);
        m->log_sd_lambda[i] = logf(m->sd_lambda[i]);
    }
    model_stats(m);
}

void nco_model_scale(const nco_model* in, const float p[6], nco_model* out) /* Pore_Model.hpp:126-138,190-201 */
{
    const float scale = p[0], shift = p[1], var = p[3], scale_sd = p[4], var_sd = p[5];
    const float log_var = logf(var), log_var_sd = logf(var_sd);
    for (unsigned i = 0; i < S; ++i)
    {
        out->level_mean[i] = in->level_mean[i] * scale + shift;
        out->level_stdv[i] = in->level_stdv[i] * var;
        out->sd_mean[i] = in->sd_mean[i] * scale_sd;
        out->sd_lambda[i] = in->sd_lambda[i] * var_sd;
        /* update_sd_stdv (:115): not read by the emission; kept for completeness */
        out->sd_stdv[i] = (float)pow(pow((double)out->sd_mean[i], 3.0) / (double)out->sd_lambda[i], .5);
        out->log_level_stdv[i] = in->log_level_stdv[i] + log_var;
        out->log_sd_lambda[i] = in->log_sd_lambda[i] + log_var_sd;
    }
    model_stats(out);
}

float nco_emission(const nco_model* m, unsigned j, float x, float y, float log_y) /* Pore_Model.hpp:24-40,145-149 */
{
    const float log_2pi = (float)log(2.0 * M_PI);
    float a = (x - m->level_mean[j]) / m->level_stdv[j];
    float ln = -m->log_level_stdv[j] - (log_2pi + a * a) / 2.0f;
    float b = (y - m->sd_mean[j]) / m->sd_mean[j];
    float li = (m->log_sd_lambda[j] - log_2pi - 3.0f * log_y - m->sd_lambda[j] * b * b / y) / 2.0f;
    return ln + li;
}

/* ------------------------------------------------------------------ State_Transitions.hpp */

unsigned nco_trans_mask(unsigned i, unsigned j)
{
    unsigned m = (i == j) ? 1u : 0u;
    for (unsigned l = 1; l < K; ++l)
        if (suffix(i, K - l) == prefix(j, K - l)) m |= 1u << l;
    return m;
}

float nco_trans_prob(unsigned i, unsigned j, float p_stay, float p_step, float p_skip_1) /* :125-144 */
{
    float p = 0;
    if (i == j) p += p_stay;
    if (suffix(i, K - 1) == prefix(j, K - 1)) p += p_step / 4;
    for (unsigned l = 2; l < K; ++l)
        if (suffix(i, K - l) == prefix(j, K - l))
            p = (float)((double)p + pow((double)p_skip_1, (double)(l - 1)) / (double)(1u << (2 * l)));
    p = (float)((double)p + (pow((double)p_skip_1, 5.0) / (double)(1.0f - p_skip_1)) / (double)S);
    return p;
}

static int cmp_u32(const void* a, const void* b)
{
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : x > y;
}

void nco_transitions(float p_stay, float p_skip, nco_trans* t) /* compute_transitions_fast :181-224 + update_fields :79-104 */
{
    float p_step = (float)(1.0 - (double)p_stay - (double)p_skip);
    float p_skip_1 = (float)((double)p_skip / ((double)p_skip + 1.0));
    memset(t->from_cnt, 0, sizeof t->from_cnt);
    for (unsigned i = 0; i < S; ++i)
    {
        /* to_s = {i} U nl1(i) U nl2(i) as an ordered set (Kmer.hpp:115-148) */
        uint32_t c[21];
        unsigned n = 0;
        c[n++] = i;
        for (unsigned b1 = 0; b1 < 4; ++b1)
        {
            unsigned i1 = (suffix(i, K - 1) << 2) + b1;
            c[n++] = i1;
            for (unsigned b2 = 0; b2 < 4; ++b2) c[n++] = (suffix(i1, K - 1) << 2) + b2;
        }
        qsort(c, n, sizeof c[0], cmp_u32);
        unsigned u = 0;
        for (unsigned k = 0; k < n; ++k)
            if (k == 0 || c[k] != c[k - 1]) c[u++] = c[k];
        t->to_cnt[i] = u;
        for (unsigned k = 0; k < u; ++k)
        {
            t->to_idx[i * NCO_MAXDEG + k] = c[k];
            t->to_lp[i * NCO_MAXDEG + k] = logf(nco_trans_prob(i, c[k], p_stay, p_step, p_skip_1));
        }
    }
    /* update_fields: from_v filled by ascending source i */
    for (unsigned i = 0; i < S; ++i)
        for (unsigned k = 0; k < t->to_cnt[i]; ++k)
        {
            unsigned j = t->to_idx[i * NCO_MAXDEG + k];
            unsigned pos = t->from_cnt[j]++;
            t->from_idx[j * NCO_MAXDEG + pos] = i;
            t->from_lp[j * NCO_MAXDEG + pos] = t->to_lp[i * NCO_MAXDEG + k];
        }
}

/* ------------------------------------------------------------------ Parameter_Trainer::init */

unsigned nco_st_train_kmers(uint32_t* out) /* Parameter_Trainer.hpp:30-57 */
{
    unsigned n = 0;
    for (unsigned i = 0; i < S; ++i)
    {
        if (nco_max_self_overlap(i) > 0) continue;
        int all_good = 1;
        for (unsigned b1 = 0; b1 < 4; ++b1)
        {
            unsigned j = (suffix(i, K - 1) << 2) + b1;
            if (nco_max_self_overlap(j) > 1) { all_good = 0; break; }
        }
        if (all_good) { if (out) out[n] = i; ++n; }
    }
    return n;
}

/* ------------------------------------------------------------------ events */

typedef struct { float* x; float* y; float* logy; const float* start; } events_t;

/* Event::update_logs (Event.hpp:33-45) + apply_drift_correction (Event.hpp:77-84) */
static int events_make(events_t* e, uint32_t n, const float* mean, const float* stdv, const float* start, float drift)
{
    e->x = (float*)malloc(sizeof(float) * (n ? n : 1));
    e->y = (float*)malloc(sizeof(float) * (n ? n : 1));
    e->logy = (float*)malloc(sizeof(float) * (n ? n : 1));
    e->start = start;
    if (!e->x || !e->y || !e->logy) return -1;
    for (uint32_t i = 0; i < n; ++i)
    {
        float s = stdv[i];
        if (s == 0.0) s = 0.01; /* double literal narrowed on store */
        e->y[i] = s;
        e->logy[i] = logf(s);
        float c = mean[i];
        c -= drift * start[i];
        e->x[i] = c;
    }
    return 0;
}
static void events_free(events_t* e) { free(e->x); free(e->y); free(e->logy); }

/* ------------------------------------------------------------------ Viterbi.hpp */

static int viterbi_core(const nco_model* pm, const nco_trans* st, const events_t* ev, uint32_t n,
                        float* path_prob, uint32_t* states, int32_t* moves,
                        float* alpha_dump, uint32_t* beta_dump)
{
    if (n == 0) return -1;
    float* alpha = (float*)malloc(sizeof(float) * (size_t)n * S);
    uint16_t* beta = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)n * S);
    if (!alpha || !beta) { free(alpha); free(beta); return -2; }
    const float log_n_states = logf((float)S);
    /* i == 0 (Viterbi.hpp:57-67) */
    for (unsigned j = 0; j < S; ++j)
    {
        alpha[j] = nco_emission(pm, j, ev->x[0], ev->y[0], ev->logy[0]) - log_n_states;
        beta[j] = S;
    }
    /* i > 0 (Viterbi.hpp:72-96): strict '>' over ascending from_v */
    for (uint32_t i = 1; i < n; ++i)
    {
        const float* prev = alpha + (size_t)(i - 1) * S;
        float* cur = alpha + (size_t)i * S;
        uint16_t* bcur = beta + (size_t)i * S;
        for (unsigned j = 0; j < S; ++j)
        {
            float best = -INFINITY;
            unsigned bp = S;
            for (unsigned k = 0; k < st->from_cnt[j]; ++k)
            {
                unsigned jp = st->from_idx[j * NCO_MAXDEG + k];
                float v = st->from_lp[j * NCO_MAXDEG + k] + prev[jp];
                if (v > best) { best = v; bp = jp; }
            }
            cur[j] = best + nco_emission(pm, j, ev->x[i], ev->y[i], ev->logy[i]);
            bcur[j] = (uint16_t)bp;
        }
    }
    /* fill_state_seq (Viterbi.hpp:120-142) */
    float max_v = -INFINITY;
    unsigned max_j = S;
    const float* last = alpha + (size_t)(n - 1) * S;
    for (unsigned j = 0; j < S; ++j)
        if (last[j] > max_v) { max_j = j; max_v = last[j]; }
    if (path_prob) *path_prob = max_v;
    int rc = 0;
    for (uint32_t i = n - 1; i > 0; --i)
    {
        if (max_j >= S) { rc = -3; break; } /* reference: out-of-range read (UB); we stop */
        if (states) states[i] = max_j;
        max_j = beta[(size_t)i * S + max_j];
    }
    if (rc == 0)
    {
        if (max_j >= S) rc = -3;
        else if (states) states[0] = max_j;
    }
    /* fill_move_seq (Viterbi.hpp:144-150) */
    if (rc == 0 && moves && states)
        for (uint32_t i = 0; i < n; ++i) moves[i] = i > 0 ? (int32_t)nco_min_skip(states[i - 1], states[i]) : 0;
    if (alpha_dump) memcpy(alpha_dump, alpha, sizeof(float) * (size_t)n * S);
    if (beta_dump) for (size_t q = 0; q < (size_t)n * S; ++q) beta_dump[q] = beta[q];
    free(alpha);
    free(beta);
    return rc;
}

/* Event_Sequence::get_base_seq (Event.hpp:85-99) */
static uint32_t base_seq(uint32_t n, const uint32_t* states, const int32_t* moves, char* out, uint32_t cap)
{
    uint32_t len = 0;
    char km[6];
    nco_kmer_string(states[0], km);
    for (unsigned c = 0; c < K; ++c) { if (out && len < cap) out[len] = km[c]; ++len; }
    for (uint32_t i = 1; i < n; ++i)
    {
        unsigned a = (unsigned)moves[i] < K ? (unsigned)moves[i] : K;
        unsigned b = K - a;
        nco_kmer_string(states[i], km);
        for (unsigned c = b; c < K; ++c) { if (out && len < cap) out[len] = km[c]; ++len; }
    }
    return len;
}

int nco_viterbi(const float* table, const float pm_params[6], float p_stay, float p_skip,
                uint32_t n, const float* mean, const float* stdv, const float* start,
                float* path_prob, uint32_t* states, int32_t* moves,
                char* bases, uint32_t bases_cap, uint32_t* n_bases,
                float* alpha_dump, uint32_t* beta_dump)
{
    if (!table || !pm_params || n == 0) return -1;
    nco_model* m0 = (nco_model*)malloc(sizeof(nco_model));
    nco_model* m = (nco_model*)malloc(sizeof(nco_model));
    nco_trans* st = (nco_trans*)malloc(sizeof(nco_trans));
    uint32_t* st_tmp = states ? NULL : (uint32_t*)malloc(sizeof(uint32_t) * n);
    int32_t* mv_tmp = moves ? NULL : (int32_t*)malloc(sizeof(int32_t) * n);
    events_t ev;
    int rc = -2;
    if (m0 && m && st && (states || st_tmp) && (moves || mv_tmp)
        && events_make(&ev, n, mean, stdv, start, pm_params[2]) == 0)
    {
        nco_model_prepare(table, m0);
        nco_model_scale(m0, pm_params, m);
        nco_transitions(p_stay, p_skip, st);
        uint32_t* sp = states ? states : st_tmp;
        int32_t* mp = moves ? moves : mv_tmp;
        rc = viterbi_core(m, st, &ev, n, path_prob, sp, mp, alpha_dump, beta_dump);
        if (rc == 0 && (bases || n_bases))
        {
            uint32_t len = base_seq(n, sp, mp, bases, bases_cap);
            if (n_bases) *n_bases = len;
        }
        events_free(&ev);
    }
    free(m0); free(m); free(st); free(st_tmp); free(mv_tmp);
    return rc;
}

typedef struct {
    const nco_model* m0; const nco_trans* st0; int all_same;
    uint32_t n_jobs; const uint64_t* ev_off;
    const float *mean, *stdv, *start, *pm_params, *st_params;
    float* path_prob; uint32_t* states; int32_t* moves;
    volatile uint32_t next; pthread_mutex_t mu; int rc;
} batch_t;

static void* batch_worker(void* arg)
{
    batch_t* b = (batch_t*)arg;
    nco_model* m = (nco_model*)malloc(sizeof(nco_model));
    nco_trans* stc = b->all_same ? NULL : (nco_trans*)malloc(sizeof(nco_trans));
    for (;;)
    {
        pthread_mutex_lock(&b->mu);
        uint32_t k = b->next < b->n_jobs ? b->next++ : UINT32_MAX;
        pthread_mutex_unlock(&b->mu);
        if (k == UINT32_MAX) break;
        const float* p = b->pm_params + 6 * k;
        uint32_t n = (uint32_t)(b->ev_off[k + 1] - b->ev_off[k]);
        uint64_t o = b->ev_off[k];
        events_t ev;
        if (events_make(&ev, n, b->mean + o, b->stdv + o, b->start + o, p[2]) != 0) { b->rc = -2; break; }
        nco_model_scale(b->m0, p, m);
        const nco_trans* st = b->st0;
        if (!b->all_same) { nco_transitions(b->st_params[2 * k], b->st_params[2 * k + 1], stc); st = stc; }
        uint32_t* sp = b->states ? b->states + o : (uint32_t*)malloc(sizeof(uint32_t) * n);
        int32_t* mp = b->moves ? b->moves + o : (int32_t*)malloc(sizeof(int32_t) * n);
        int rc = viterbi_core(m, st, &ev, n, b->path_prob ? b->path_prob + k : NULL, sp, mp, NULL, NULL);
        if (rc != 0) b->rc = rc;
        if (!b->states) free(sp);
        if (!b->moves) free(mp);
        events_free(&ev);
    }
    free(m); free(stc);
    return NULL;
}

int nco_viterbi_batch(const float* table, uint32_t n_jobs, const uint64_t* ev_off,
                      const float* mean, const float* stdv, const float* start,
                      const float* pm_params, const float* st_params, uint32_t n_threads,
                      float* path_prob, uint32_t* states, int32_t* moves)
{
    if (n_jobs == 0) return 0;
    nco_model* m0 = (nco_model*)malloc(sizeof(nco_model));
    nco_trans* st0 = (nco_trans*)malloc(sizeof(nco_trans));
    if (!m0 || !st0) { free(m0); free(st0); return -2; }
    nco_model_prepare(table, m0);
    batch_t b;
    memset(&b, 0, sizeof b);
    b.all_same = 1;
    for (uint32_t k = 1; k < n_jobs; ++k)
        if (st_params[2 * k] != st_params[0] || st_params[2 * k + 1] != st_params[1]) b.all_same = 0;
    if (b.all_same) nco_transitions(st_params[0], st_params[1], st0);
    b.m0 = m0; b.st0 = st0; b.n_jobs = n_jobs; b.ev_off = ev_off;
    b.mean = mean; b.stdv = stdv; b.start = start; b.pm_params = pm_params; b.st_params = st_params;
    b.path_prob = path_prob; b.states = states; b.moves = moves;
    pthread_mutex_init(&b.mu, NULL);
    if (n_threads <= 1) batch_worker(&b);
    else
    {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
        for (uint32_t t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, batch_worker, &b);
        for (uint32_t t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    pthread_mutex_destroy(&b.mu);
    free(m0); free(st0);
    return b.rc;
}

/* ------------------------------------------------------------------ Forward_Backward.hpp */

typedef struct { uint32_t n; float* alpha; float* beta; float log_pr_data; } fwbw_t;

static int fwbw_core(const nco_model* pm, const nco_trans* st, const events_t* ev, uint32_t n, fwbw_t* fb)
{
    fb->n = n;
    fb->alpha = (float*)malloc(sizeof(float) * (size_t)n * S);
    fb->beta = (float*)calloc((size_t)n * S, sizeof(float));
    float* emit = (float*)malloc(sizeof(float) * S);
    if (!fb->alpha || !fb->beta || !emit) { free(emit); return -2; }
    const float log_n_states = logf((float)S);
    for (unsigned j = 0; j < S; ++j) /* :58-70 */
        fb->alpha[j] = nco_emission(pm, j, ev->x[0], ev->y[0], ev->logy[0]) - log_n_states;
    for (uint32_t i = 1; i < n; ++i) /* :74-89 */
    {
        const float* prev = fb->alpha + (size_t)(i - 1) * S;
        float* cur = fb->alpha + (size_t)i * S;
        for (unsigned j = 0; j < S; ++j)
        {
            float s = -INFINITY;
            for (unsigned k = 0; k < st->from_cnt[j]; ++k)
                s = nco_flogsum(s, st->from_lp[j * NCO_MAXDEG + k] + prev[st->from_idx[j * NCO_MAXDEG + k]]);
            cur[j] = nco_emission(pm, j, ev->x[i], ev->y[i], ev->logy[i]) + s;
        }
    }
    /* beta[n-1] = 0 (:93-103); beta[i] for i < n-1 (:107-125).  The reference recomputes the
     * emission per edge; it is the same value per (successor, event), so compute it once. */
    for (uint32_t ip1 = n - 1; ip1 > 0; --ip1)
    {
        uint32_t i = ip1 - 1;
        const float* nxt = fb->beta + (size_t)ip1 * S;
        float* cur = fb->beta + (size_t)i * S;
        for (unsigned j = 0; j < S; ++j) emit[j] = nco_emission(pm, j, ev->x[ip1], ev->y[ip1], ev->logy[ip1]);
        for (unsigned j = 0; j < S; ++j)
        {
            float s = -INFINITY;
            for (unsigned k = 0; k < st->to_cnt[j]; ++k)
            {
                unsigned jn = st->to_idx[j * NCO_MAXDEG + k];
                s = nco_flogsum(s, st->to_lp[j * NCO_MAXDEG + k] + emit[jn] + nxt[jn]);
            }
            cur[j] += s; /* zero-initialised, so this is an assignment (:120) */
        }
    }
    float s = -INFINITY; /* :129-134 */
    for (unsigned j = 0; j < S; ++j) s = nco_flogsum(s, fb->alpha[(size_t)(n - 1) * S + j]);
    fb->log_pr_data = s;
    free(emit);
    return 0;
}
static void fwbw_free(fwbw_t* fb) { free(fb->alpha); free(fb->beta); fb->alpha = fb->beta = NULL; }

int nco_fwbw(const float* table, const float pm_params[6], float p_stay, float p_skip,
             uint32_t n, const float* mean, const float* stdv, const float* start,
             float* alpha, float* beta, float* log_pr_data)
{
    if (!table || !pm_params || n == 0) return -1;
    nco_model* m0 = (nco_model*)malloc(sizeof(nco_model));
    nco_model* m = (nco_model*)malloc(sizeof(nco_model));
    nco_trans* st = (nco_trans*)malloc(sizeof(nco_trans));
    events_t ev;
    fwbw_t fb = { 0, NULL, NULL, 0 };
    int rc = -2;
    if (m0 && m && st && events_make(&ev, n, mean, stdv, start, pm_params[2]) == 0)
    {
        nco_model_prepare(table, m0);
        nco_model_scale(m0, pm_params, m);
        nco_transitions(p_stay, p_skip, st);
        rc = fwbw_core(m, st, &ev, n, &fb);
        if (rc == 0)
        {
            if (alpha) memcpy(alpha, fb.alpha, sizeof(float) * (size_t)n * S);
            if (beta) memcpy(beta, fb.beta, sizeof(float) * (size_t)n * S);
            if (log_pr_data) *log_pr_data = fb.log_pr_data;
        }
        fwbw_free(&fb);
        events_free(&ev);
    }
    free(m0); free(m); free(st);
    return rc;
}

/* ------------------------------------------------------------------ Parameter_Trainer.hpp */

#define MAXSEQ 16

typedef struct {
    uint32_t n_seqs;
    uint32_t len[MAXSEQ], strand[MAXSEQ];
    const float *mean[MAXSEQ], *stdv[MAXSEQ], *start[MAXSEQ]; /* uncorrected events */
    events_t cev[MAXSEQ];                                       /* drift-corrected copies */
    fwbw_t fb[MAXSEQ];
    const nco_model* model[2];   /* unscaled */
    nco_model* scaled[2];
    nco_trans* trans[2];
    float fit;
} train_data;

static float log_posterior(const fwbw_t* fb, uint32_t i, unsigned j) /* Forward_Backward.hpp:41 */
{
    return fb->alpha[(size_t)i * S + j] + fb->beta[(size_t)i * S + j] - fb->log_pr_data;
}

/* train_pm_params (Parameter_Trainer.hpp:230-427) */
static void train_pm_params(const train_data* d, const float crt[6], int train_drift, float out[6], int* done)
{
    *done = 0;
    unsigned total_n_events = 0;
    double A[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
    double B[3] = { 0, 0, 0 };
    double D = 0.0, V_numer = 0.0, V_denom = 0.0, U_pos = 0.0;
    for (unsigned k = 0; k < d->n_seqs; ++k)
    {
        unsigned st = d->strand[k];
        uint32_t n_events = d->len[k];
        total_n_events += n_events;
        const nco_model* pm = d->model[st];
        const fwbw_t* fb = &d->fb[k];
        for (uint32_t i = 0; i < n_events; ++i)
        {
            float x_i = d->mean[k][i];
            float y_i = d->cev[k].y[i]; /* events[i].stdv after update_logs' 0 -> 0.01 */
            float t_i = d->start[k][i];
            float s[3] = { 0, 0, 0 }, l[3] = { 0, 0, 0 };
            for (unsigned j = 0; j < S; ++j)
            {
                float p_ij = expf(log_posterior(fb, i, j));
                float term_s0 = p_ij / (pm->level_stdv[j] * pm->level_stdv[j]);
                float term_s1 = term_s0 * pm->level_mean[j];
                float term_s2 = term_s1 * pm->level_mean[j];
                float term_l0 = p_ij * pm->sd_lambda[j];
                float term_l1 = term_l0 / pm->sd_mean[j];
                float term_l2 = term_l1 / pm->sd_mean[j];
                s[0] += term_s0; s[1] += term_s1; s[2] += term_s2;
                l[0] += term_l0; l[1] += term_l1; l[2] += term_l2;
            }
            A[0][0] += s[0];
            A[0][1] += s[1];
            A[1][1] += s[2];
            B[0] += s[0] * x_i;
            B[1] += s[1] * x_i;
            if (train_drift)
            {
                A[0][2] += s[0] * t_i;
                A[1][2] += s[1] * t_i;
                A[2][2] += s[0] * t_i * t_i;
                B[2] += s[0] * x_i * t_i;
            }
            D += s[0] * x_i * x_i;
            V_numer += l[2] * y_i;
            V_denom += l[1];
            U_pos += l[0] / y_i;
        }
    }
    A[1][0] = A[0][1];
    A[2][0] = A[0][2];
    A[2][1] = A[1][2];
    if (!train_drift) A[2][2] = 1.0;
    double Ac[3][3], Bc[3], C[3];
    memcpy(Ac, A, sizeof A);
    memcpy(Bc, B, sizeof B);
    for (unsigned i = 0; i < 3; ++i) /* alg::max_value_of: first maximum under '<' */
    {
        C[i] = A[i][0];
        for (unsigned j = 1; j < 3; ++j) if (C[i] < A[i][j]) C[i] = A[i][j];
    }
    for (unsigned i = 0; i < 3; ++i)
    {
        unsigned p = i;
        double p_val = fabs(A[i][i]) / C[p];
        for (unsigned i2 = i + 1; i2 < 3; ++i2)
        {
            double i2_val = fabs(A[i2][i]) / C[i2];
            if (i2_val > p_val) { p = i2; p_val = i2_val; }
        }
        if (p_val < 1e-7)
        {
            *done = 1;
            memcpy(out, crt, 6 * sizeof(float));
            return;
        }
        if (p > i)
        {
            double tmp;
            for (unsigned j = 0; j < 3; ++j) { tmp = A[i][j]; A[i][j] = A[p][j]; A[p][j] = tmp; }
            tmp = B[i]; B[i] = B[p]; B[p] = tmp;
            tmp = C[i]; C[i] = C[p]; C[p] = tmp;
        }
        for (p = i + 1; p < 3; ++p)
        {
            double m = A[p][i] / A[i][i];
            A[p][i] = 0;
            for (unsigned j = i + 1; j < 3; ++j) A[p][j] -= m * A[i][j];
            B[p] -= m * B[i];
        }
    }
    /* a_hat = shift, b_hat = scale, c_hat = drift, d_hat = var, v_hat = scale_sd, u_hat = var_sd;
     * each is a float lvalue, so every assignment narrows (:236-241) */
    float c_hat = (float)(B[2] / A[2][2]);
    float b_hat = (float)((B[1] - A[1][2] * c_hat) / A[1][1]);
    float a_hat = (float)((B[0] - A[0][1] * b_hat - A[0][2] * c_hat) / A[0][0]);
    double d_numer = (D
                      + a_hat * a_hat * Ac[0][0]
                      + b_hat * b_hat * Ac[1][1]
                      + c_hat * c_hat * Ac[2][2]
                      + 2.0 * a_hat * b_hat * Ac[0][1]
                      + 2.0 * a_hat * c_hat * Ac[0][2]
                      + 2.0 * b_hat * c_hat * Ac[1][2]
                      - 2.0 * (a_hat * Bc[0]
                               + b_hat * Bc[1]
                               + c_hat * Bc[2]));
    float d_hat = (float)sqrt(d_numer / (double)total_n_events);
    float v_hat = (float)(V_numer / V_denom);
    float u_hat = (float)((double)total_n_events / (U_pos - V_denom / v_hat));
    out[0] = b_hat; out[1] = a_hat; out[2] = c_hat; out[3] = d_hat; out[4] = v_hat; out[5] = u_hat;
}

/* train_st_params (Parameter_Trainer.hpp:434-532) */
static void train_st_params(const train_data* d, const float crt_st[4], float out_st[4])
{
    static uint32_t kmers[S];
    static unsigned n_kmers = 0;
    if (n_kmers == 0) n_kmers = nco_st_train_kmers(kmers);
    for (unsigned st = 0; st < 2; ++st)
    {
        float s_stay = -INFINITY, s_skip = -INFINITY, s_denom = -INFINITY;
        float p_stay = crt_st[2 * st], p_skip = crt_st[2 * st + 1];
        float log_p_stay = logf(p_stay);
        float log_p_step_4 = (float)(log(1.0 - (double)p_stay - (double)p_skip) - log(4.0));
        for (unsigned k = 0; k < d->n_seqs; ++k)
        {
            if (d->strand[k] != st) continue;
            const nco_model* spm = d->scaled[st];
            const events_t* ce = &d->cev[k];
            const fwbw_t* fb = &d->fb[k];
            uint32_t n_events = d->len[k];
#define JOINT(i, j1, j2, lt) \
    (fb->alpha[(size_t)(i) * S + (j1)] + (lt) \
     + nco_emission(spm, (j2), ce->x[(i) + 1], ce->y[(i) + 1], ce->logy[(i) + 1]) \
     + fb->beta[(size_t)((i) + 1) * S + (j2)] - fb->log_pr_data)
            for (uint32_t i = 0; i + 1 < n_events; ++i)
            {
                for (unsigned q = 0; q < n_kmers; ++q)
                {
                    unsigned j1 = kmers[q];
                    float log_p_j1 = log_posterior(fb, i, j1);
                    s_denom = nco_flogsum(s_denom, log_p_j1);
                    float log_p_j1_j1 = JOINT(i, j1, j1, log_p_stay);
                    if (log_p_j1_j1 > log_p_j1) log_p_j1_j1 = log_p_j1;
                    s_stay = nco_flogsum(s_stay, log_p_j1_j1);
                    float s2 = -INFINITY;
                    s2 = nco_flogsum(s2, log_p_j1_j1);
                    for (unsigned b1 = 0; b1 < 4; ++b1)
                    {
                        unsigned j2 = (suffix(j1, K - 1) << 2) + b1;
                        s2 = nco_flogsum(s2, JOINT(i, j1, j2, log_p_step_4));
                    }
                    float log_p_j1_d01 = s2;
                    if (log_p_j1_d01 > log_p_j1) log_p_j1_d01 = log_p_j1;
                    float p_j1_d2 = expf(log_p_j1) - expf(log_p_j1_d01);
                    s_skip = nco_flogsum(s_skip, logf(p_j1_d2));
                }
            }
#undef JOINT
        }
        float n_stay = expf(s_stay - s_denom);
        float n_skip = expf(s_skip - s_denom);
        if (n_stay < .05 || n_stay > .4 || n_skip < .05 || n_skip > .4) /* compared as double (:518-519) */
        {
            float a_stay = n_stay > .05f ? n_stay : .05f; /* std::max(x, .05f) */
            a_stay = .4f < a_stay ? .4f : a_stay;          /* std::min(x, .4f) */
            float a_skip = n_skip > .05f ? n_skip : .05f;
            a_skip = .4f < a_skip ? .4f : a_skip;
            n_stay = a_stay;
            n_skip = a_skip;
        }
        out_st[2 * st] = n_stay;
        out_st[2 * st + 1] = n_skip;
    }
}

int nco_train_one_round(uint32_t n_seqs, const uint32_t* seq_len, const uint32_t* seq_strand,
                        const float* mean, const float* stdv, const float* start,
                        const float* table0, const float* table1,
                        const float pm_params[6], const float st_params[4],
                        int train_scaling, int train_transitions, int train_drift,
                        float new_pm_params[6], float new_st_params[4], float* fit, int* done)
{
    if (n_seqs == 0 || n_seqs > MAXSEQ) return -1;
    train_data d;
    memset(&d, 0, sizeof d);
    nco_model* um[2] = { (nco_model*)malloc(sizeof(nco_model)), (nco_model*)malloc(sizeof(nco_model)) };
    nco_model_prepare(table0, um[0]);
    nco_model_prepare(table1, um[1]);
    d.model[0] = um[0];
    d.model[1] = um[1];
    d.n_seqs = n_seqs;
    size_t off = 0;
    int rc = 0;
    /* fill_train_data (Parameter_Trainer.hpp:99-155) */
    for (unsigned k = 0; k < n_seqs; ++k)
    {
        unsigned st = seq_strand[k];
        if (st > 1) { rc = -1; break; }
        d.len[k] = seq_len[k];
        d.strand[k] = st;
        d.mean[k] = mean + off; d.stdv[k] = stdv + off; d.start[k] = start + off;
        off += seq_len[k];
        if (!d.scaled[st])
        {
            d.scaled[st] = (nco_model*)malloc(sizeof(nco_model));
            nco_model_scale(d.model[st], pm_params, d.scaled[st]);
            d.trans[st] = (nco_trans*)malloc(sizeof(nco_trans));
            nco_transitions(st_params[2 * st], st_params[2 * st + 1], d.trans[st]);
        }
    }
    d.fit = 0.0f;
    unsigned made = 0;
    for (unsigned k = 0; rc == 0 && k < n_seqs; ++k)
    {
        unsigned st = d.strand[k];
        if (events_make(&d.cev[k], d.len[k], d.mean[k], d.stdv[k], d.start[k], pm_params[2]) != 0) { rc = -2; break; }
        rc = fwbw_core(d.scaled[st], d.trans[st], &d.cev[k], d.len[k], &d.fb[k]);
        ++made;
        if (rc == 0) d.fit += d.fb[k].log_pr_data;
    }
    if (rc == 0)
    {
        /* train_one_round (:541-579).  The caller's new_* start as copies of crt_* (nanocall.cpp:369-380). */
        memcpy(new_pm_params, pm_params, 6 * sizeof(float));
        memcpy(new_st_params, st_params, 4 * sizeof(float));
        *fit = d.fit;
        *done = 0;
        int dn = 0;
        if (train_scaling)
        {
            train_pm_params(&d, pm_params, train_drift, new_pm_params, &dn);
            *done = dn;
        }
        if (!dn && train_transitions) train_st_params(&d, st_params, new_st_params);
    }
    for (unsigned k = 0; k < made; ++k) { fwbw_free(&d.fb[k]); events_free(&d.cev[k]); }
    for (unsigned st = 0; st < 2; ++st) { free(d.scaled[st]); free(d.trans[st]); free(um[st]); }
    return rc;
}
