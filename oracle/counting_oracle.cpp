/*
 * counting_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product) for the counting paths:
 * declare mining (existences, ordered relations, positions), the pair-index view of a CSR log and the
 * trace-id intersection.  Restates the reference's Spark jobs literally, one trace at a time, over the
 * "SeqTable view" of the pair index: a trace is listed under pair (A,B) iff it holds an A
 * before a B (for A == B: at least two occurrences).
 *
 * Parity pins: the reference has NO tests for com.datalab.siesta.queryprocessor.declare, so
 * these restatements are pinned by code reading only; tests/test_counting_oracle.py checks them against
 * hand-computed examples and algebraic identities.
 *
 * J/ = src/main/java/com/datalab/siesta/queryprocessor/
 */
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {

/* Number of int64 values of the packed result (same layout as the product's siesta_declare_counts_device). */
int64_t oracle_declare_size(int32_t A, int32_t k_cap) {
    return 4ll * A + (int64_t)A * (k_cap + 1) + 8ll * A * A + 2;
}

/*
 * Packed layout (int64): tot[A] uniq[A] first[A] last[A] hist[A][k_cap+1] co[A][A] ordered[A][A]
 *                        response[A][A] precedence[A][A] alt_response[A][A] alt_precedence[A][A]
 *                        chain_response[A][A] chain_precedence[A][A] hist_overflow n_nonempty
 *
 *  tot[a]        total occurrences of a                  (S3Connector.querySingleTableDeclare :317-335 summed in
 *                                                         QueryPlanOrderedRelations.execute :73-77)
 *  uniq[a]       traces containing a                     (QueryPlanExistences.extractUniqueTracesSingle :150-158)
 *  hist[a][k]    traces with exactly k occurrences, k>=1 (UniqueTracesPerEventType.groupTimes :31-41)
 *  co[a][b]      |traces(a,b) U traces(b,a)|             (QueryPlanExistences.joinUnionTraces :164-179)
 *  ordered[a][b] traces listed under pair (a,b)          (S3Connector.queryIndexTableDeclare :389-404); > 0 <=> the
 *                                                         key exists (DeclareUtilities.extractNotFoundPairs :25-42)
 *  response[a][b]   sum over traces listed under (a,b), a != b, of #{a : exists b after a}
 *                   (OrderedRelationsUtilityFunctions.countResponse :25-31 via joinTables :97-116)
 *  precedence[a][b] same with #{b : exists a before b}   (countPrecedence :38-44)
 *  alt_response / alt_precedence / chain_response / chain_precedence: the same sums with
 *                   countResponseAlternate :51-60, countPrecedenceAlternate :67-76, countResponseChain :83-89,
 *                   countPrecedenceChain :96-102 (QueryPlanOrderedRelationsAlternate / ...Chain.evaluateConstraint)
 *  first[a]/last[a] traces whose first/last event is a   (QueryPlanPositions.execute :51-79)
 */
int oracle_declare_counts(const int64_t* trace_off, const int32_t* act, int64_t n_traces, int32_t A, int32_t k_cap,
                          int64_t* out) {
    const int64_t n_out = oracle_declare_size(A, k_cap);
    std::memset(out, 0, sizeof(int64_t) * (size_t)n_out);
    int64_t* tot = out;
    int64_t* uniq = tot + A;
    int64_t* first = uniq + A;
    int64_t* last = first + A;
    int64_t* hist = last + A;
    int64_t* co = hist + (int64_t)A * (k_cap + 1);
    int64_t* ordered = co + (int64_t)A * A;
    int64_t* response = ordered + (int64_t)A * A;
    int64_t* precedence = response + (int64_t)A * A;
    int64_t* alt_r = precedence + (int64_t)A * A;
    int64_t* alt_p = alt_r + (int64_t)A * A;
    int64_t* chain_r = alt_p + (int64_t)A * A;
    int64_t* chain_p = chain_r + (int64_t)A * A;
    int64_t* hist_overflow = chain_p + (int64_t)A * A;
    int64_t* n_nonempty = hist_overflow + 1;

    std::vector<std::vector<int>> pos((size_t)A);
    for (int64_t t = 0; t < n_traces; ++t) {
        const int64_t lo = trace_off[t], hi = trace_off[t + 1];
        if (hi <= lo) continue;
        ++*n_nonempty;
        for (auto& p : pos) p.clear();
        for (int64_t i = lo; i < hi; ++i)
            if (act[i] >= 0 && act[i] < A) pos[act[i]].push_back((int)(i - lo));
        if (act[lo] >= 0 && act[lo] < A) ++first[act[lo]];
        if (act[hi - 1] >= 0 && act[hi - 1] < A) ++last[act[hi - 1]];
        for (int a = 0; a < A; ++a) {
            const int64_t c = (int64_t)pos[a].size();
            if (c == 0) continue;
            tot[a] += c;
            ++uniq[a];
            if (c <= k_cap) ++hist[(int64_t)a * (k_cap + 1) + c];
            else ++*hist_overflow;
        }
        for (int a = 0; a < A; ++a) {
            if (pos[a].empty()) continue;
            for (int b = 0; b < A; ++b) {
                if (pos[b].empty()) continue;
                /* listed under (a,b): an a strictly before a b */
                bool ab = false, ba = false;
                for (int x : pos[a])
                    for (int y : pos[b]) {
                        if (x < y) ab = true;
                        if (y < x) ba = true;
                    }
                if (ab) ++ordered[(int64_t)a * A + b];
                if (ab || ba) ++co[(int64_t)a * A + b];
                if (ab && a != b) { /* QueryPlanOrderedRelations.execute filters eventA != eventB (:62) */
                    int64_t r = 0, p = 0;
                    for (int x : pos[a]) {
                        bool any = false;
                        for (int y : pos[b]) any = any || y > x;
                        r += any;
                    }
                    for (int y : pos[b]) {
                        bool any = false;
                        for (int x : pos[a]) any = any || x < y;
                        p += any;
                    }
                    response[(int64_t)a * A + b] += r;
                    precedence[(int64_t)a * A + b] += p;
                    /* literal restatements of the alternate / chain counters; pos[] is ascending = the "sorted" lists */
                    const std::vector<int>& la = pos[a];
                    const std::vector<int>& lb = pos[b];
                    int64_t ar = 0, ap = 0, cr = 0, cp = 0;
                    for (size_t i = 0; i + 1 < la.size(); ++i) {
                        bool any = false;
                        for (int y : lb) any = any || (y > la[i] && y < la[i + 1]);
                        ar += any;
                    }
                    {
                        bool any = false;
                        for (int y : lb) any = any || y > la.back();
                        ar += any;
                    }
                    for (size_t i = 1; i < lb.size(); ++i) {
                        bool any = false;
                        for (int y : la) any = any || (y < lb[i] && y > lb[i - 1]);
                        ap += any;
                    }
                    {
                        bool any = false;
                        for (int y : la) any = any || y < lb[0];
                        ap += any;
                    }
                    for (int x : la) {
                        bool any = false;
                        for (int y : lb) any = any || y == x + 1;
                        cr += any;
                    }
                    for (int y : lb) {
                        bool any = false;
                        for (int x : la) any = any || x == y - 1;
                        cp += any;
                    }
                    alt_r[(int64_t)a * A + b] += ar;
                    alt_p[(int64_t)a * A + b] += ap;
                    chain_r[(int64_t)a * A + b] += cr;
                    chain_p[(int64_t)a * A + b] += cp;
                }
            }
        }
    }
    return 0;
}

/* Posting list of pair (a,b) under the SeqTable view: ascending trace indices.  Returns the length. */
int64_t oracle_posting_list(const int64_t* trace_off, const int32_t* act, int64_t n_traces, int32_t a, int32_t b,
                            int64_t* out) {
    int64_t n = 0;
    for (int64_t t = 0; t < n_traces; ++t) {
        bool seen_a = false, hit = false;
        for (int64_t i = trace_off[t]; i < trace_off[t + 1] && !hit; ++i) {
            if (seen_a && act[i] == b) hit = true;
            if (act[i] == a) seen_a = true;
        }
        if (hit) out[n++] = t;
    }
    return n;
}

/* SparkDatabaseRepository.getCommonIds (J/storage/repositories/SparkDatabaseRepository.java:160-178): the traces
 * that appear in EVERY one of the `n_lists` posting lists (lists are ascending and duplicate-free).  Returns the
 * number of common ids written to out (ascending). */
int64_t oracle_intersect(const int64_t* const* lists, const int64_t* lens, int32_t n_lists, int64_t* out) {
    if (n_lists <= 0) return 0;
    std::vector<int64_t> cur(lists[0], lists[0] + lens[0]);
    for (int k = 1; k < n_lists; ++k) {
        std::vector<int64_t> nxt;
        std::set_intersection(cur.begin(), cur.end(), lists[k], lists[k] + lens[k], std::back_inserter(nxt));
        cur.swap(nxt);
    }
    std::copy(cur.begin(), cur.end(), out);
    return (int64_t)cur.size();
}

/*
 * Pair statistics behind /stats.  PARITY UNPINNED: the reference only READS these numbers from count.parquet
 * (QueryPlanStats.execute, J/model/Queries/QueryPlans/QueryPlanStats.java:43-48; J/model/DBModel/Count.java:12-24);
 * the table is written by the SIESTA preprocess component, which is not in the reference repository.  This restates
 * the pairing policy SIESTA publishes for its index — non-overlapping skip-till-next-match pairs: per trace and pair
 * (a,b) take the first a, the first b after it, emit (a,b), continue after that b — so the oracle and the GPU kernel
 * agree on a stated definition, not on a reference run.
 * out[p] = count, sum, min, max (0 if none), sum of squares low / high 64 bits.
 */
int oracle_pair_stats(const int64_t* trace_off, const int32_t* act, const int64_t* ts_ms, int64_t n_traces,
                      const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs, int64_t* out) {
    for (int p = 0; p < n_pairs; ++p) {
        int64_t cnt = 0, sum = 0, mn = 0, mx = 0;
        unsigned __int128 sq = 0;
        for (int64_t t = 0; t < n_traces; ++t) {
            bool waiting_b = false;
            int64_t ta = 0;
            for (int64_t i = trace_off[t]; i < trace_off[t + 1]; ++i) {
                if (waiting_b) {
                    if (act[i] == pair_b[p]) {
                        const int64_t d = ts_ms[i] - ta;
                        mn = cnt == 0 ? d : std::min(mn, d);
                        mx = cnt == 0 ? d : std::max(mx, d);
                        ++cnt;
                        sum += d;
                        sq += (unsigned __int128)((__int128)d * (__int128)d);
                        waiting_b = false;
                    }
                } else if (act[i] == pair_a[p]) {
                    waiting_b = true;
                    ta = ts_ms[i];
                }
            }
        }
        int64_t* o = out + (int64_t)p * 6;
        o[0] = cnt; o[1] = sum; o[2] = mn; o[3] = mx;
        o[4] = (int64_t)(uint64_t)sq;
        o[5] = (int64_t)(uint64_t)(sq >> 64);
    }
    return 0;
}

}  /* extern "C" */
