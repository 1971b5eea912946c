/* sc_oracle.c -- plain-C CPU oracle for the gym_supplychain step/reset dynamics.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may build,
 * load or call this file (through oracle/c_oracle.py).  The product (gym_supplychain_b200) never does.
 *
 * It restates -- not copies -- the reference algorithm, one env at a time, in IEEE doubles with every addition in
 * the reference's order (paths relative to the reference repository root):
 *
 *   SC_Action.apply (SUPPLY / SHIP cuts) .... gym_supplychain/envs/supplychain_env.py:49-96
 *   SC_Node.act ............................. supplychain_env.py:208-396
 *   SC_Node.reset ........................... supplychain_env.py:402-412
 *   SC_Node.build_observation ............... supplychain_env.py:428-463
 *   SupplyChainEnv.step / _build_observation  supplychain_env.py:697-748, 762-791
 *   BeerGameEnv.step / reset ................ gym_supplychain/envs/beergame_env.py:66-156
 *   BeerGameEnv2.step / reset ............... gym_supplychain/envs/beergame2_env.py:79-192
 *
 * In-transit material is a binary heap of (time, amount) laid out exactly like CPython's heapq list
 * (Lib/heapq.py: heappush/_siftdown, heappop/_siftup), because the reference observation reads that list by index
 * (SURVEY.md F3).  Stochastic inputs (demand rows, lead times) are always INJECTED by the caller: the reference draws
 * them from numpy's MT19937, which stays in Python (oracle/sc_oracle.py).
 *
 * Pinned by tests/test_c_oracle.py: bit-for-bit against every golden vector recorded from the unmodified reference
 * (tests/golden/*.npz: observations, rewards, done flags, stocks, raw heap arrays) and against the Python oracle on
 * random specs.  Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/c_oracle.py); no fused multiply-adds, so
 * every product and sum rounds like numpy's.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct ScoSpec {
    int32_t N, P, E, T, avgL, maxL, stochL, n_init;
    double c_unmet, c_xstock, c_xproc, c_xship;
    const int32_t* last;      /* [N] */
    const int32_t* ebeg;      /* [N+1] edges of node n: ebeg[n] .. ebeg[n+1]-1, in `destinations` order */
    const double* proc_cap;   /* [N] */
    const double *stock0, *cap, *hold, *sup_cap, *sup_cost, *ratio, *proc_cost; /* [N*P] */
    const int32_t* edst;      /* [E] */
    const double* ecap;       /* [E] */
    const double* ecost;      /* [E*P]  dest_costs[p][i] at (ebeg[n]+i)*P + p */
    const int32_t* ibeg;      /* [N*P+1] initial in-transit entries of (n, p), in push order */
    const int32_t* itime;     /* [n_init] arrival period (1-based) */
    const double* iamt;       /* [n_init] */
    const double *dlo, *dspan; /* [P] demand normalisation of the observation (:770-777) */
} ScoSpec;

typedef struct { int len, cap; int* tm; double* am; } Heap;

typedef struct Sco {
    ScoSpec s;
    int R, A, O, K;
    int* is_sup; int* can_ship; double* max_ship; int* n_supply; int* n_ship; int* ret_idx;
    double* stock; Heap* q;
    const int32_t* demand;    /* [T+1][R][P], caller-owned */
    const int32_t* lts;       /* [T][K] or NULL */
    int t;
} Sco;

/* ---- CPython heapq on (time, amount) tuples ------------------------------------------------- */
static int lt_item(int ta, double aa, int tb, double ab) { return ta < tb || (ta == tb && aa < ab); }

static void heap_reserve(Heap* h, int need) {
    if (need <= h->cap) return;
    int c = h->cap ? h->cap * 2 : 8;
    while (c < need) c *= 2;
    h->tm = (int*)realloc(h->tm, (size_t)c * sizeof(int));
    h->am = (double*)realloc(h->am, (size_t)c * sizeof(double));
    h->cap = c;
}
static void sift_down(Heap* h, int start, int pos) {            /* heapq._siftdown */
    const int nt = h->tm[pos]; const double na = h->am[pos];
    while (pos > start) {
        const int par = (pos - 1) >> 1;
        if (lt_item(nt, na, h->tm[par], h->am[par])) { h->tm[pos] = h->tm[par]; h->am[pos] = h->am[par]; pos = par; }
        else break;
    }
    h->tm[pos] = nt; h->am[pos] = na;
}
static void sift_up(Heap* h, int pos) {                         /* heapq._siftup */
    const int end = h->len, start = pos;
    const int nt = h->tm[pos]; const double na = h->am[pos];
    int child = 2 * pos + 1;
    while (child < end) {
        const int right = child + 1;
        if (right < end && !lt_item(h->tm[child], h->am[child], h->tm[right], h->am[right])) child = right;
        h->tm[pos] = h->tm[child]; h->am[pos] = h->am[child];
        pos = child; child = 2 * pos + 1;
    }
    h->tm[pos] = nt; h->am[pos] = na;
    sift_down(h, start, pos);
}
static void heap_push(Heap* h, int when, double amount) {
    heap_reserve(h, h->len + 1);
    h->tm[h->len] = when; h->am[h->len] = amount; h->len += 1;
    sift_down(h, 0, h->len - 1);
}
static double heap_pop(Heap* h) {                               /* returns the amount of the smallest item */
    h->len -= 1;
    const int lt_ = h->tm[h->len]; const double la = h->am[h->len];
    if (h->len == 0) return la;
    const double ret = h->am[0];
    h->tm[0] = lt_; h->am[0] = la;
    sift_up(h, 0);
    return ret;
}

/* ---- construction --------------------------------------------------------------------------- */
static void* dup_mem(const void* p, size_t bytes) { void* q = malloc(bytes ? bytes : 1); if (bytes) memcpy(q, p, bytes); return q; }

Sco* sco_create(const ScoSpec* sp) {
    Sco* o = (Sco*)calloc(1, sizeof(Sco));
    o->s = *sp;
    const int N = sp->N, P = sp->P, NP = N * P, E = sp->E;
    ScoSpec* s = &o->s;
    s->last = (const int32_t*)dup_mem(sp->last, (size_t)N * 4); s->ebeg = (const int32_t*)dup_mem(sp->ebeg, (size_t)(N + 1) * 4);
    s->proc_cap = (const double*)dup_mem(sp->proc_cap, (size_t)N * 8);
    s->stock0 = (const double*)dup_mem(sp->stock0, (size_t)NP * 8); s->cap = (const double*)dup_mem(sp->cap, (size_t)NP * 8);
    s->hold = (const double*)dup_mem(sp->hold, (size_t)NP * 8); s->sup_cap = (const double*)dup_mem(sp->sup_cap, (size_t)NP * 8);
    s->sup_cost = (const double*)dup_mem(sp->sup_cost, (size_t)NP * 8); s->ratio = (const double*)dup_mem(sp->ratio, (size_t)NP * 8);
    s->proc_cost = (const double*)dup_mem(sp->proc_cost, (size_t)NP * 8);
    s->edst = (const int32_t*)dup_mem(sp->edst, (size_t)E * 4); s->ecap = (const double*)dup_mem(sp->ecap, (size_t)E * 8);
    s->ecost = (const double*)dup_mem(sp->ecost, (size_t)E * P * 8);
    s->ibeg = (const int32_t*)dup_mem(sp->ibeg, (size_t)(NP + 1) * 4);
    s->itime = (const int32_t*)dup_mem(sp->itime, (size_t)sp->n_init * 4); s->iamt = (const double*)dup_mem(sp->iamt, (size_t)sp->n_init * 8);
    s->dlo = (const double*)dup_mem(sp->dlo, (size_t)P * 8); s->dspan = (const double*)dup_mem(sp->dspan, (size_t)P * 8);
    o->is_sup = (int*)calloc((size_t)N, sizeof(int)); o->can_ship = (int*)calloc((size_t)NP, sizeof(int));
    o->max_ship = (double*)calloc((size_t)NP, sizeof(double)); o->n_supply = (int*)calloc((size_t)N, sizeof(int));
    o->n_ship = (int*)calloc((size_t)N, sizeof(int)); o->ret_idx = (int*)calloc((size_t)N, sizeof(int));
    o->stock = (double*)calloc((size_t)NP, sizeof(double)); o->q = (Heap*)calloc((size_t)NP, sizeof(Heap));
    for (int n = 0; n < N; ++n) {                                /* :137, :147-152 */
        double mx = 0;
        for (int p = 0; p < P; ++p) if (s->sup_cap[n * P + p] > mx) mx = s->sup_cap[n * P + p];
        o->is_sup[n] = mx > 0;
        for (int p = 0; p < P; ++p) o->max_ship[n * P + p] = o->is_sup[n] ? s->sup_cap[n * P + p] : 0.0;
    }
    for (int n = 0; n < N; ++n) {                                /* :196-206 */
        const int D = s->ebeg[n + 1] - s->ebeg[n];
        if (D > 0) {
            for (int p = 0; p < P; ++p) o->can_ship[n * P + p] = s->cap[n * P + p] > 0;
            for (int i = 0; i < D; ++i)
                for (int p = 0; p < P; ++p) o->max_ship[s->edst[s->ebeg[n] + i] * P + p] += s->ecap[s->ebeg[n] + i];
        }
    }
    int R = 0, A = 0, K = 0;
    for (int n = 0; n < N; ++n) {                                /* :597-610 */
        const int D = s->ebeg[n + 1] - s->ebeg[n];
        int ns = 0, nsh = 0;
        if (o->is_sup[n]) for (int p = 0; p < P; ++p) if (s->sup_cap[n * P + p] > 0) ++ns;
        for (int p = 0; p < P; ++p) if (o->can_ship[n * P + p]) nsh += D;
        o->n_supply[n] = ns; o->n_ship[n] = nsh;
        A += ns + nsh;
        K += (ns > 0 ? P : 0) + D;
        o->ret_idx[n] = s->last[n] ? R++ : -1;
    }
    o->R = R; o->A = A; o->K = K;
    o->O = R * P + N * P * (1 + s->avgL) + 1;                    /* :617-621 */
    return o;
}

void sco_destroy(Sco* o) {
    if (!o) return;
    ScoSpec* s = &o->s;
    free((void*)s->last); free((void*)s->ebeg); free((void*)s->proc_cap); free((void*)s->stock0); free((void*)s->cap);
    free((void*)s->hold); free((void*)s->sup_cap); free((void*)s->sup_cost); free((void*)s->ratio); free((void*)s->proc_cost);
    free((void*)s->edst); free((void*)s->ecap); free((void*)s->ecost); free((void*)s->ibeg); free((void*)s->itime);
    free((void*)s->iamt); free((void*)s->dlo); free((void*)s->dspan);
    for (int i = 0; i < s->N * s->P; ++i) { free(o->q[i].tm); free(o->q[i].am); }
    free(o->is_sup); free(o->can_ship); free(o->max_ship); free(o->n_supply); free(o->n_ship); free(o->ret_idx);
    free(o->stock); free(o->q); free(o);
}

int sco_n_actions(const Sco* o) { return o->A; }
int sco_n_obs(const Sco* o) { return o->O; }
int sco_n_leadtimes(const Sco* o) { return o->K; }
int sco_n_retailers(const Sco* o) { return o->R; }

/* ---- reset (:402-412, :630-682) --------------------------------------------------------------- */
void sco_reset(Sco* o, const int32_t* demand, const int32_t* lts) {
    const ScoSpec* s = &o->s;
    o->demand = demand; o->lts = lts; o->t = 0;
    for (int np = 0; np < s->N * s->P; ++np) {
        o->stock[np] = s->stock0[np];
        o->q[np].len = 0;
        for (int k = s->ibeg[np]; k < s->ibeg[np + 1]; ++k) heap_push(&o->q[np], s->itime[k], s->iamt[k]);
    }
}

/* ---- SHIP cuts (:58-96): ascending (value, index); each destination gets its slice of `limit`, never more than is left */
static void ship_cuts(const double* v, int D, double limit, double* out, int* order) {
    for (int i = 0; i < D; ++i) {                                /* stable insertion sort == sorted((value, index)) */
        int j = i;
        while (j > 0 && v[order[j - 1]] > v[i]) { order[j] = order[j - 1]; --j; }
        order[j] = i;
    }
    double left = limit, prev = 0;
    for (int k = 0; k < D; ++k) {
        const int i = order[k];
        double amount = (v[i] - prev) * limit;
        if (amount > left) amount = left;
        out[i] = amount;
        left -= amount;
        prev = v[i];
    }
}

/* ---- one node, one period (:208-396) ---------------------------------------------------------- */
static double node_step(Sco* o, int n, const double* acts, const int* lts, int t, const int32_t* dem) {
    const ScoSpec* s = &o->s;
    const int P = s->P, D = s->ebeg[n + 1] - s->ebeg[n], e0 = s->ebeg[n];
    double cost = 0;
    for (int p = 0; p < P; ++p) {                                /* arrivals, ascending (time, amount) (:220-228) */
        Heap* q = &o->q[n * P + p];
        double got = 0.0;
        while (q->len > 0 && q->tm[0] == t) got += heap_pop(q);
        o->stock[n * P + p] = o->stock[n * P + p] + got;
    }
    for (int p = 0; p < P; ++p) {                                /* over capacity: pay and discard (:232-240) */
        const int np = n * P + p;
        if (o->stock[np] > s->cap[np]) { cost += s->c_xstock * (o->stock[np] - s->cap[np]); o->stock[np] = s->cap[np]; }
    }
    int ai = 0, li = 0;
    if (o->is_sup[n] && o->n_supply[n] > 0) {                     /* supply (:244-256); cursor moves only when supplied (F4b) */
        for (int p = 0; p < P; ++p) {
            const int np = n * P + p;
            if (s->sup_cap[np] > 0) {
                const double amount = acts[ai] * s->sup_cap[np];
                ++ai;
                if (amount > 0) { heap_push(&o->q[np], t + lts[li], amount); ++li; }
                cost += amount * s->sup_cost[np];
            }
        }
    }
    if (!s->last[n]) {                                            /* ship (:262-375) */
        double room[64], out[64], send[64]; int order[64];
        for (int i = 0; i < D; ++i) room[i] = s->ecap[e0 + i];
        double proc_left = s->proc_cap[n];
        const int factory = s->proc_cap[n] > 0;
        const int base = li;
        for (int p = 0; p < P; ++p) {
            const int np = n * P + p;
            if (!o->can_ship[np]) continue;
            double over_ship = 0, over_proc = 0;
            if (o->stock[np] > 0) {
                const double limit = s->cap[np] < o->stock[np] ? s->cap[np] : o->stock[np];   /* min(capacity, stock) (:61-65) */
                ship_cuts(acts + ai, D, limit, out, order);
                for (int i = 0; i < D; ++i) send[i] = out[i];
                if (factory) {                                     /* :298-310 */
                    for (int i = 0; i < D; ++i) {
                        if (out[i] > 0) {
                            if (out[i] > proc_left) { over_proc += out[i] - proc_left; out[i] = proc_left; }
                            proc_left -= out[i];
                        }
                        send[i] = out[i] / s->ratio[np];
                    }
                }
                for (int i = 0; i < D; ++i) {                      /* :312-328, capacity only touched on a clip, in raw units (F4a) */
                    if (send[i] > 0 && send[i] > room[i]) {
                        over_ship += send[i] - room[i];
                        send[i] = room[i];
                        out[i] = factory ? send[i] * s->ratio[np] : send[i];
                        room[i] -= out[i];
                    }
                }
                double leaving = 0;
                for (int i = 0; i < D; ++i) leaving = leaving + out[i];            /* :331 */
                o->stock[np] -= leaving;
                if (factory) cost += leaving * s->proc_cost[np];                   /* :338 */
                for (int i = 0; i < D; ++i)                                         /* :344-348 */
                    if (send[i] > 0) heap_push(&o->q[s->edst[e0 + i] * P + p], t + lts[base + i], send[i]);
                double sc = 0;
                for (int i = 0; i < D; ++i) sc = sc + send[i] * s->ecost[(e0 + i) * P + p];   /* :352-353 */
                cost += sc;
            }
            cost += s->c_xproc * over_proc;                        /* :361 */
            cost += s->c_xship * over_ship;                        /* :366 */
            ai += D;
        }
    } else {                                                       /* retailers (:380-384) */
        for (int p = 0; p < P; ++p) {
            const int np = n * P + p;
            const double d = (double)dem[p];
            const double served = o->stock[np] < d ? o->stock[np] : d;
            o->stock[np] -= served;
            cost += s->c_unmet * (d - served);
        }
    }
    for (int p = 0; p < P; ++p) cost += o->stock[n * P + p] * s->hold[n * P + p];   /* :390-391 */
    return cost;
}

/* ---- step (:703-748): returns done; *reward = -sum of node costs -------------------------------- */
int sco_step(Sco* o, const double* action, double* reward) {
    const ScoSpec* s = &o->s;
    const int P = s->P;
    double a[4096]; int lt[4096];
    for (int i = 0; i < o->A; ++i) a[i] = (action[i] + 1) / 2;     /* :697-698 */
    o->t += 1;
    const int t = o->t;
    double total = 0;
    int ai = 0, li = 0, ri = 0;
    for (int n = 0; n < s->N; ++n) {
        const int na = o->n_supply[n] + o->n_ship[n];
        int nl;
        if (s->stochL) {                                           /* :719-722 */
            nl = o->n_supply[n] + o->n_ship[n] / P;
            for (int k = 0; k < nl; ++k) lt[k] = o->lts[(size_t)(t - 1) * o->K + li + k];
            li += nl;
        } else {
            nl = na;
            for (int k = 0; k < nl; ++k) lt[k] = s->avgL;          /* :724 */
        }
        const int32_t* dem = NULL;
        if (s->last[n]) { dem = o->demand + ((size_t)(t - 1) * o->R + ri) * P; ++ri; }   /* :728 */
        total += node_step(o, n, a + ai, lt, t, dem);
        ai += na;
    }
    *reward = -total;
    return t == s->T;
}

/* ---- observation (:428-463, :762-791).  canonical = 0: the reference layout (walks each heap ARRAY by index, F3);
 *      canonical = 1: true per-arrival-time sums in ascending (time, amount) order (what the fast GPU modes emit). */
static int cmp_item(const void* a, const void* b) {
    const double* x = (const double*)a; const double* y = (const double*)b;      /* (time as double, amount) */
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1] ? 1 : 0);
}
void sco_observation(const Sco* o, double* obs, int canonical) {
    const ScoSpec* s = &o->s;
    const int P = s->P, L = s->avgL, t = o->t;
    int k = 0;
    for (int r = 0; r < o->R; ++r)
        for (int p = 0; p < P; ++p) obs[k++] = ((double)o->demand[((size_t)t * o->R + r) * P + p] - s->dlo[p]) / s->dspan[p];
    for (int n = 0; n < s->N; ++n) {
        for (int p = 0; p < P; ++p) obs[k++] = o->stock[n * P + p] / s->cap[n * P + p];
        for (int p = 0; p < P; ++p) {
            const Heap* q = &o->q[n * P + p];
            const double ms = o->max_ship[n * P + p];
            if (q->len == 0) { for (int i = 0; i < L; ++i) obs[k++] = 0; continue; }
            if (!canonical) {
                int idx = 0;
                for (int when = t + 1; when < t + L; ++when) {
                    double sum = 0;
                    while (idx < q->len && q->tm[idx] == when) { sum = sum + q->am[idx]; ++idx; }
                    obs[k++] = sum / ms;
                }
                double sum = 0;
                while (idx < q->len) { sum = sum + q->am[idx]; ++idx; }
                obs[k++] = sum / (ms * (double)(s->maxL - (L - 1)));
            } else {
                double* it = (double*)malloc((size_t)q->len * 2 * sizeof(double));
                for (int i = 0; i < q->len; ++i) { it[2 * i] = (double)q->tm[i]; it[2 * i + 1] = q->am[i]; }
                qsort(it, (size_t)q->len, 2 * sizeof(double), cmp_item);
                for (int when = t + 1; when < t + L; ++when) {
                    double sum = 0;
                    for (int i = 0; i < q->len; ++i) if ((int)it[2 * i] == when) sum = sum + it[2 * i + 1];
                    obs[k++] = sum / ms;
                }
                double sum = 0;
                for (int i = 0; i < q->len; ++i) if ((int)it[2 * i] >= t + L) sum = sum + it[2 * i + 1];
                obs[k++] = sum / (ms * (double)(s->maxL - (L - 1)));
                free(it);
            }
        }
    }
    obs[k++] = (double)(s->T - t) / (double)s->T;
    for (int i = 0; i < k; ++i) {                                  /* clip(2x - 1, -1, 1) (:700-701) */
        double x = obs[i] * 2 - 1;
        obs[i] = x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);
    }
}

void sco_stocks(const Sco* o, double* out) { memcpy(out, o->stock, (size_t)o->s.N * o->s.P * sizeof(double)); }
int sco_heap(const Sco* o, int np, int cap, int32_t* tm, double* am) {
    const Heap* q = &o->q[np];
    for (int i = 0; i < q->len && i < cap; ++i) { tm[i] = q->tm[i]; am[i] = q->am[i]; }
    return q->len;
}

/* One whole rollout of `T` periods (the agent loop of tests/utils.py:13-22) with the obs of every period. */
void sco_rollout(Sco* o, int T, const double* actions, double* obs, double* reward, unsigned char* done, int canonical) {
    for (int k = 0; k < T; ++k) {
        done[k] = (unsigned char)sco_step(o, actions + (size_t)k * o->A, reward + k);
        if (obs) sco_observation(o, obs + (size_t)k * o->O, canonical);
    }
}

/* =================================================================================================
 * Beer game (all integers).  version 0: beergame_env.py:66-156 (action = delta on the incoming order);
 * version 2: beergame2_env.py:79-192 (action = the order, capacity penalty).  demand[T] and delays[T] (weeks 1..T)
 * are injected; the delay in force for the initial pipeline is 2 (beergame_env.py:39,52).
 * ================================================================================================= */
typedef struct Bgo {
    int version, L, T, inv_cost, backlog_cost, max_stock, penalty, init_ship, init_orders;
    int inv0[16], inv[16], backlog[16], placed[16];
    long long* transit;        /* [T + max_delay + 3][L] by absolute arrival week */
    int horizon;
    const int32_t *demand, *delays;
    int week;
} Bgo;

Bgo* bgo_create(int version, int L, int T, int inv_cost, int backlog_cost, int max_stock, int penalty, int init_ship,
                int init_orders, const int32_t* inv0) {
    Bgo* b = (Bgo*)calloc(1, sizeof(Bgo));
    b->version = version; b->L = L; b->T = T; b->inv_cost = inv_cost; b->backlog_cost = backlog_cost; b->max_stock = max_stock;
    b->penalty = penalty; b->init_ship = init_ship; b->init_orders = init_orders;
    for (int i = 0; i < L; ++i) b->inv0[i] = inv0[i];
    b->horizon = T + 64;
    b->transit = (long long*)calloc((size_t)b->horizon * L, sizeof(long long));
    return b;
}
void bgo_destroy(Bgo* b) { if (b) { free(b->transit); free(b); } }

void bgo_obs(const Bgo* b, long long* obs) {                       /* beergame2_env.py:111-112 */
    const int off = b->version == 2 ? b->max_stock : 0;
    for (int i = 0; i < b->L; ++i) obs[i] = (long long)off + b->inv[i] - b->backlog[i];
}
void bgo_reset(Bgo* b, const int32_t* demand, const int32_t* delays) {
    b->demand = demand; b->delays = delays; b->week = 0;
    memset(b->transit, 0, (size_t)b->horizon * b->L * sizeof(long long));
    for (int i = 0; i < b->L; ++i) { b->inv[i] = b->inv0[i]; b->backlog[i] = 0; b->placed[i] = b->init_orders; }
    for (int w = 1; w <= 2; ++w) for (int i = 0; i < b->L; ++i) b->transit[(size_t)w * b->L + i] = b->init_ship;   /* :52 */
}
int bgo_step(Bgo* b, const int32_t* action, long long* reward) {
    const int L = b->L;
    b->week += 1;
    const int w = b->week;
    int incoming[16], to_fill[16], deliver[16];
    for (int i = 0; i < L; ++i) { b->inv[i] += (int)b->transit[(size_t)w * L + i]; b->transit[(size_t)w * L + i] = 0; }   /* :72 */
    incoming[0] = b->demand[w - 1];                                /* :79-81 */
    for (int i = 1; i < L; ++i) incoming[i] = b->placed[i - 1];
    for (int i = 0; i < L; ++i) {
        to_fill[i] = incoming[i] + b->backlog[i];
        deliver[i] = b->inv[i] < to_fill[i] ? b->inv[i] : to_fill[i];
    }
    const int delay = b->delays[w - 1];
    if (delay == 0) { for (int i = 0; i + 1 < L; ++i) b->inv[i] += deliver[i + 1]; }          /* :93-96 */
    else { for (int i = 0; i + 1 < L; ++i) b->transit[(size_t)(w + delay) * L + i] += deliver[i + 1]; }
    for (int i = 0; i < L; ++i) { b->inv[i] -= deliver[i]; b->backlog[i] = to_fill[i] - deliver[i]; }   /* :101-103 */
    if (delay == 0) b->inv[L - 1] += b->placed[L - 1];             /* :111-114 */
    else b->transit[(size_t)(w + delay) * L + L - 1] += b->placed[L - 1];
    for (int i = 0; i < L; ++i) b->placed[i] = b->version == 0 ? incoming[i] + action[i] : action[i];   /* :121 / v2 :168 */
    long long r = 0;
    for (int i = 0; i < L; ++i) r += (long long)b->inv_cost * b->inv[i] + (long long)b->backlog_cost * b->backlog[i];
    if (b->version == 2)                                           /* beergame2_env.py:179-180 */
        for (int i = 0; i < L; ++i) {
            const int xi = b->inv[i] - b->max_stock, xb = b->backlog[i] - b->max_stock;
            r += (long long)b->penalty * ((xi > 0 ? xi : 0) + (xb > 0 ? xb : 0));
        }
    *reward = -r;
    return w == b->T;
}
