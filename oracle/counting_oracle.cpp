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
    return 4ll * A + (int64_t)A * (k_cap + 1) + 4ll * A * A + 2;
}

/*
 * Packed layout (int64): tot[A] uniq[A] first[A] last[A] hist[A][k_cap+1] co[A][A] ordered[A][A]
 *                        response[A][A] precedence[A][A] hist_overflow n_nonempty
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
    int64_t* hist_overflow = precedence + (int64_t)A * A;
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

}  /* extern "C" */
