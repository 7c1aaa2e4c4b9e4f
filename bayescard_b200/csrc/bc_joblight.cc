// Native planner for star joins over the IMDB ensemble (BASELINE.json config 3): a BATCH of job-light SQL
// texts -> the reference's factor lists in array form, off the Python hot loop (the Python planner + per-dict decoding cost
// ~8 us per factor against < 0.1 us of kernel time).
//
// Restates bayescard_b200/joblight.py, which restates what Evaluation/parse_query_imdb.py:54-325 (DeepDB's generate_factors /
// factor_refine over two-table models, rdc-based choice of the first model) produces for the only shape job-light has:
// a star on title.id over the five models title x X (Schemas/imdb/schema.py:59-63).  For a query over title and tables
// a, b, ... with conditions C_t, C_a, C_b, ...:
//
//   card = |J_a| * E_a[ 1{C_t, C_a, a not null} * prod_{b != a} F_b ]                       first model a (largest pairwise-RDC vector)
//          * prod_{b != a, C_b not empty}  P_b(C_b, C_t, b not null) / P_b(C_t, b not null)  nominator, then denominator (inverse)
//
// Factors of one query are consecutive; BN_ensemble.cardinality's combination rule (Models/BN_ensemble_model.py:228-252) is
// bc_joblight_combine.  Predicates are emitted against the column indices of each model's bc_sqlc (bc_sqlc_column_index), ready
// for bc_sqlc_compile_factors.  NOT PINNED against the reference planner (it cannot run: SURVEY.md section 8c); pinned to the
// Python mirror row for row (tests/test_joblight.py) and, through it, to the 70 shipped true cardinalities.
#include <algorithm>
#include <cctype>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "bc_internal.h"

using sv = std::string_view;

struct bc_joblight {
    int n_bn = 0;
    std::vector<const bc_sqlc*> sqlc;        // per BN: the column tables (not owned)
    std::vector<std::string> table;          // per BN: the table X of title x X
    std::vector<double> join_size;
    std::vector<int32_t> fan_node;           // [a * n_bn + b]: node of title.mul_<X_b>.movie_id in BN a, -1 if absent
    std::map<std::string, int> rdc_id;       // "table.column" names that have a pairwise RDC entry -> row of rdc_mat
    std::vector<double> rdc_mat;             // [n_rdc_names x n_rdc_names], NaN = no entry (both orders filled)
    int n_rdc_names = 0;
    std::vector<int32_t> nn_col;             // [bn * n_bn + table]: column of "<table>.<table>_nn" in model bn
    double epsilon = 0.1;
    std::map<std::string, std::string> alias;  // default alias -> table (the FROM clause overrides)
};

namespace {

// A conditioned column, interned per planning thread: (table, column name) -> its full "table.column" name, its place in the RDC
// table and its column index in every model (resolved on first use).  A workload touches a dozen of these; linear search.
struct ColEnt {
    int table;                 // -1: title, else BN index
    std::string col, full;
    int gid;                   // row of bc_joblight::rdc_mat, -1 if the name has no RDC entry
    std::vector<int> bn_col;   // per BN: bc_sqlc_column_index(full), INT_MIN = not looked up yet
};

struct Cond {
    int table;
    int col;          // index into Scratch::cols
    int op;           // 0 '=', 1 '<', 2 '>', 3 '<=', 4 '>='
    double val;
};

struct Factor {
    int32_t bn;
    uint8_t inverse;
    uint32_t fan_mask;
    uint32_t n_pred;
};

struct ColQ { int col; bool has_eq; double eq, lo, hi; };

struct Plan {   // per thread
    std::vector<Factor> factors;
    std::vector<int32_t> pcol;
    std::vector<uint8_t> pkind;
    std::vector<double> pa, pb;
    std::vector<uint32_t> q_nfactors;   // per query
    std::vector<uint8_t> q_status;      // 0 ok, 1 not a job-light star query (left to the Python planner)
    std::vector<double> q_join;
    // scratch, reused across queries (no allocation in steady state)
    std::vector<ColEnt> cols;
    std::vector<Cond> conds;
    std::vector<ColQ> c_t, c_x;
    std::vector<int> names;
    std::vector<const ColQ*> vals;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
inline char lower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }
sv strip(sv s) {
    while (!s.empty() && is_space(s.front())) s.remove_prefix(1);
    while (!s.empty() && is_space(s.back())) s.remove_suffix(1);
    return s;
}
bool ieq(sv a, const char* b) {
    const size_t n = std::strlen(b);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; ++i)
        if (lower(a[i]) != lower(b[i])) return false;
    return true;
}
// case-insensitive search of a (lower-case) keyword surrounded by whitespace
size_t find_kw(sv s, const char* kw, size_t from = 0) {
    const size_t n = std::strlen(kw);
    const char k0 = kw[0];
    for (size_t i = from; i + n <= s.size(); ++i) {
        if (lower(s[i]) != k0) continue;
        bool ok = true;
        for (size_t j = 1; j < n && ok; ++j) ok = lower(s[i + j]) == kw[j];
        if (ok && (i == 0 || is_space(s[i - 1])) && (i + n == s.size() || is_space(s[i + n]))) return i;
    }
    return sv::npos;
}
bool is_ident(sv s) {
    if (s.empty()) return false;
    for (char ch : s)
        if (!((ch >= '0' && ch <= '9') || (ch >= 'a' && ch <= 'z') || (ch >= 'A' && ch <= 'Z') || ch == '_')) return false;
    return true;
}

int intern_col(const bc_joblight& h, Plan& p, int table, sv col) {
    for (size_t i = 0; i < p.cols.size(); ++i)
        if (p.cols[i].table == table && p.cols[i].col == col) return (int)i;
    ColEnt e;
    e.table = table;
    e.col = std::string(col);
    e.full = (table < 0 ? std::string("title") : h.table[table]) + "." + e.col;
    auto it = h.rdc_id.find(e.full);
    e.gid = it == h.rdc_id.end() ? -1 : it->second;
    e.bn_col.assign(h.n_bn, INT_MIN);
    p.cols.push_back(std::move(e));
    return (int)p.cols.size() - 1;
}
inline int col_in_bn(const bc_joblight& h, Plan& p, int col, int bn) {
    int& v = p.cols[col].bn_col[bn];
    if (v == INT_MIN) v = bc_sqlc_column_index(h.sqlc[bn], p.cols[col].full.c_str());
    return v;
}

// columns of one table in first-mention order, each {eq | (lo, hi)}  (joblight.py _table_query)
void table_query(const bc_joblight& h, const std::vector<Cond>& conds, int table, std::vector<ColQ>& q) {
    q.clear();
    for (const Cond& c : conds) {
        if (c.table != table) continue;
        ColQ* cq = nullptr;
        for (ColQ& x : q)
            if (x.col == c.col) cq = &x;
        if (!cq) { q.push_back(ColQ{c.col, false, 0, -HUGE_VAL, HUGE_VAL}); cq = &q.back(); }
        switch (c.op) {
            case 0: cq->has_eq = true; cq->eq = c.val; break;
            case 2: cq->lo = std::max(cq->lo, c.val + h.epsilon); break;
            case 4: cq->lo = std::max(cq->lo, c.val); break;
            case 1: cq->hi = std::min(cq->hi, c.val - h.epsilon); break;
            case 3: cq->hi = std::min(cq->hi, c.val); break;
        }
    }
}

// One SQL text -> factors appended to `out`.  Returns false when the text is not a job-light star query.
bool plan_one(const bc_joblight& h, sv sql, Plan& out) {
    sql = strip(sql);
    while (!sql.empty() && sql.back() == ';') sql.remove_suffix(1);
    const size_t f = find_kw(sql, "from"), w = find_kw(sql, "where");
    if (f == sv::npos || w == sv::npos || w < f) return false;
    {   // "select count(*)"
        sv head = strip(sql.substr(0, f));
        if (head.size() < 6 || !ieq(head.substr(0, 6), "select")) return false;
    }
    // ---- FROM: "table alias, table alias, ..."
    struct Alias { sv name; int id; };   // alias -> table id (-1 title, BN index otherwise); a later duplicate replaces the earlier one
    Alias alias_of[16];
    int n_alias = 0;
    int order[16], n_order = 0;          // joined tables (BN index) in FROM order
    bool has_title = false;
    {
        sv from = sql.substr(f + 4, w - f - 4);
        size_t pos = 0;
        while (pos <= from.size()) {
            size_t comma = from.find(',', pos);
            if (comma == sv::npos) comma = from.size();
            sv part = strip(from.substr(pos, comma - pos));
            pos = comma + 1;
            if (part.empty()) return false;
            size_t sp = 0;
            while (sp < part.size() && !is_space(part[sp])) ++sp;
            const sv table = part.substr(0, sp);
            sv al = strip(part.substr(sp));
            if (al.size() > 3 && ieq(al.substr(0, 2), "as") && is_space(al[2])) al = strip(al.substr(3));   // "table AS alias"
            {   // the last token is the alias (joblight.py: toks[-1])
                size_t last = al.find_last_of(" \t");
                if (last != sv::npos) al = al.substr(last + 1);
            }
            if (al.empty()) al = table;
            int id = -2;
            if (table == "title") { id = -1; has_title = true; }
            else
                for (int b = 0; b < h.n_bn; ++b)
                    if (table == h.table[b]) id = b;
            if (id == -2) return false;   // a table outside the star
            int slot = -1;
            for (int i = 0; i < n_alias; ++i)
                if (alias_of[i].name == al) slot = i;
            if (slot < 0) {
                if (n_alias == 16) return false;
                slot = n_alias++;
            }
            alias_of[slot] = Alias{al, id};
            if (id >= 0) {
                if (n_order == 16) return false;
                order[n_order++] = id;
            }
            if (comma == from.size()) break;
        }
    }
    if (!has_title || n_order == 0) return false;
    // ---- WHERE: conditions "alias.col op number" joined by AND; join conditions "a.x = b.y" are skipped
    std::vector<Cond>& conds = out.conds;
    conds.clear();
    {
        sv where = sql.substr(w + 5);
        size_t pos = 0;
        while (pos < where.size()) {
            size_t nx = find_kw(where, "and", pos);
            sv c = strip(where.substr(pos, nx == sv::npos ? sv::npos : nx - pos));
            pos = nx == sv::npos ? where.size() : nx + 3;
            if (c.empty()) return false;
            size_t o = c.find_first_of("<>=");
            if (o == sv::npos) return false;
            size_t o2 = o + 1;
            if (o2 < c.size() && c[o2] == '=') ++o2;
            const sv lhs = strip(c.substr(0, o)), ops = c.substr(o, o2 - o), rhs = strip(c.substr(o2));
            const size_t dot = lhs.find('.');
            if (dot == sv::npos || !is_ident(lhs.substr(0, dot)) || !is_ident(lhs.substr(dot + 1))) return false;
            {   // join condition: the right-hand side is alias.column
                const size_t rdot = rhs.find('.');
                // (the mirror tests its join pattern \w+.\w+ = \w+.\w+ first, so "t.x = 2005.5" is skipped like a join condition)
                if (ops == "=" && rdot != sv::npos && is_ident(rhs.substr(0, rdot)) && is_ident(rhs.substr(rdot + 1))) continue;
            }
            int table = -2;
            {
                const sv al = lhs.substr(0, dot);
                for (int i = 0; i < n_alias; ++i)
                    if (alias_of[i].name == al) table = alias_of[i].id;
            }
            if (table == -2) return false;
            // number: -?digits(.digits)?
            {
                size_t i = 0;
                if (i < rhs.size() && rhs[i] == '-') ++i;
                size_t d0 = i;
                while (i < rhs.size() && rhs[i] >= '0' && rhs[i] <= '9') ++i;
                if (i == d0) return false;
                if (i < rhs.size() && rhs[i] == '.') {
                    ++i;
                    size_t d1 = i;
                    while (i < rhs.size() && rhs[i] >= '0' && rhs[i] <= '9') ++i;
                    if (i == d1) return false;
                }
                if (i != rhs.size()) return false;
            }
            Cond cd;
            cd.table = table;
            cd.col = intern_col(h, out, table, lhs.substr(dot + 1));
            cd.op = ops == "=" ? 0 : ops == "<" ? 1 : ops == ">" ? 2 : ops == "<=" ? 3 : ops == ">=" ? 4 : -1;
            if (cd.op < 0) return false;
            {   // correctly rounded, like float() in the mirror
                char buf[64];
                if (rhs.size() >= sizeof(buf)) return false;
                std::memcpy(buf, rhs.data(), rhs.size());
                buf[rhs.size()] = 0;
                cd.val = std::strtod(buf, nullptr);
            }
            conds.push_back(cd);
        }
    }
    table_query(h, conds, -1, out.c_t);
    const std::vector<ColQ>& c_t = out.c_t;
    // ---- first model: _greedily_select_first_cardinality_spn with rdc_spn_selection (joblight.py _first_table)
    int first = -1;
    {
        int sorted[16], n_sorted = n_order;
        std::copy(order, order + n_order, sorted);
        std::sort(sorted, sorted + n_sorted);
        n_sorted = (int)(std::unique(sorted, sorted + n_sorted) - sorted);
        double best_rdc = 0;
        int best_where = 0;
        for (int si = 0; si < n_sorted; ++si) {
            const int t = sorted[si];
            int cols[64], n_cols = 0;   // set of conditioned columns over {title, t}, first-mention order (title first)
            for (int tb : {-1, t})
                for (const Cond& c : conds)
                    if (c.table == tb && std::find(cols, cols + n_cols, c.col) == cols + n_cols && n_cols < 64) cols[n_cols++] = c.col;
            double rdc = 0;
            for (int xi = 0; xi < n_cols; ++xi)
                for (int yi = 0; yi < n_cols; ++yi) {
                    const ColEnt &x = out.cols[cols[xi]], &y = out.cols[cols[yi]];
                    if (x.gid >= 0 && y.gid >= 0 && x.full < y.full) {
                        const double v = h.rdc_mat[(size_t)x.gid * h.n_rdc_names + y.gid];
                        if (!std::isnan(v)) rdc += v;
                    }
                }
            int n_where = 0;
            for (int tb : {-1, t}) {
                bool any = false;
                for (const Cond& c : conds) any = any || c.table == tb;
                n_where += any;
            }
            if (first < 0 || rdc > best_rdc || (rdc == best_rdc && n_where > best_where)) {
                first = t;
                best_rdc = rdc;
                best_where = n_where;
            }
        }
    }
    // ---- factors
    auto emit = [&](int bn, bool inverse, uint32_t fan_mask, const std::vector<ColQ>* part0, const std::vector<ColQ>* part1, int nn_table) {
        Factor fc{bn, (uint8_t)inverse, fan_mask, 0};
        // dict semantics of the Python planner: a later update() of the same key replaces the value but keeps the position
        // (a key is a (table, column) pair = one interned column)
        out.names.clear();
        out.vals.clear();
        for (const std::vector<ColQ>* part : {part0, part1}) {
            if (!part) continue;
            for (const ColQ& cq : *part) {
                size_t j = 0;
                for (; j < out.names.size(); ++j)
                    if (out.names[j] == cq.col) break;
                if (j == out.names.size()) { out.names.push_back(cq.col); out.vals.push_back(&cq); }
                else out.vals[j] = &cq;
            }
        }
        for (size_t j = 0; j < out.names.size(); ++j) {
            out.pcol.push_back(col_in_bn(h, out, out.names[j], bn));   // -1: KeyError in the mirror -> the factor compiler flags the factor
            out.pkind.push_back(out.vals[j]->has_eq ? 0 : 1);
            out.pa.push_back(out.vals[j]->has_eq ? out.vals[j]->eq : out.vals[j]->lo);
            out.pb.push_back(out.vals[j]->has_eq ? 0.0 : out.vals[j]->hi);
            ++fc.n_pred;
        }
        // the NOT NULL condition of relevant_conditions: <table>.<table>_nn = 1
        out.pcol.push_back(h.nn_col[(size_t)bn * h.n_bn + nn_table]);
        out.pkind.push_back(0);
        out.pa.push_back(1.0);
        out.pb.push_back(0.0);
        ++fc.n_pred;
        out.factors.push_back(fc);
    };
    uint32_t nf = 0;
    {
        table_query(h, conds, first, out.c_x);
        uint32_t mask = 0;
        for (int i = 0; i < n_order; ++i)
            if (order[i] != first) {
                const int node = h.fan_node[(size_t)first * h.n_bn + order[i]];
                if (node < 0 || node >= 32) return false;   // no such fan-out column in the first model: leave it to Python
                mask |= 1u << node;
            }
        emit(first, false, mask, &c_t, &out.c_x, first);
        ++nf;
    }
    uint32_t seen = 0;
    for (int i = 0; i < n_order; ++i) {
        const int b = order[i];
        if (b == first || ((seen >> b) & 1u)) continue;
        seen |= 1u << b;
        table_query(h, conds, b, out.c_x);
        if (out.c_x.empty()) continue;   // factor_refine: nominator and denominator cancel
        emit(b, false, 0, &c_t, &out.c_x, b);
        emit(b, true, 0, &c_t, nullptr, b);
        nf += 2;
    }
    out.q_nfactors.back() = nf;
    out.q_join.back() = h.join_size[first];
    return true;
}

}  // namespace

extern "C" {

int bc_joblight_create(int n_bn, const bc_sqlc* const* sqlc, const char* const* tables, const double* join_sizes, const int32_t* fan_node,
                       int n_rdc, const char* const* rdc_a, const char* const* rdc_b, const double* rdc_val, double epsilon, bc_joblight** out) {
    if (!out || n_bn <= 0 || n_bn > 32 || !sqlc || !tables || !join_sizes || !fan_node) { bc_set_error("bc_joblight_create: bad arguments"); return BC_EINVAL; }
    bc_joblight* h = new bc_joblight();
    h->n_bn = n_bn;
    for (int b = 0; b < n_bn; ++b) {
        if (!sqlc[b] || !tables[b]) { delete h; bc_set_error("bc_joblight_create: model %d is NULL", b); return BC_EINVAL; }
        h->sqlc.push_back(sqlc[b]);
        h->table.emplace_back(tables[b]);
        h->join_size.push_back(join_sizes[b]);
    }
    h->fan_node.assign(fan_node, fan_node + (size_t)n_bn * n_bn);
    for (int i = 0; i < n_rdc; ++i)
        for (const char* nm : {rdc_a[i], rdc_b[i]})
            if (!h->rdc_id.count(nm)) { const int id = (int)h->rdc_id.size(); h->rdc_id[nm] = id; }
    h->n_rdc_names = (int)h->rdc_id.size();
    h->rdc_mat.assign((size_t)h->n_rdc_names * h->n_rdc_names, std::nan(""));
    for (int i = 0; i < n_rdc; ++i) {
        const int a = h->rdc_id[rdc_a[i]], b = h->rdc_id[rdc_b[i]];
        h->rdc_mat[(size_t)a * h->n_rdc_names + b] = rdc_val[i];
        h->rdc_mat[(size_t)b * h->n_rdc_names + a] = rdc_val[i];
    }
    h->nn_col.assign((size_t)n_bn * n_bn, -1);
    for (int bn = 0; bn < n_bn; ++bn)
        for (int t = 0; t < n_bn; ++t) {
            const std::string nm = h->table[t] + "." + h->table[t] + "_nn";
            h->nn_col[(size_t)bn * n_bn + t] = bc_sqlc_column_index(h->sqlc[bn], nm.c_str());
        }
    h->epsilon = epsilon;
    *out = h;
    return BC_OK;
}

void bc_joblight_destroy(bc_joblight* h) { delete h; }

}  // extern "C"

namespace {

// Plans queries [0, n) (text of query q = text_of(q)) on host threads; the per-thread results are copied into the caller's arrays
// by the same threads (offsets by prefix sums), so nothing of the batch is serial but the sums.
template <class TextOf>
int plan_batch(const bc_joblight* h, size_t n, TextOf text_of, uint8_t* status, double* join_size, uint32_t* first_factor, size_t factor_capacity,
               int32_t* factor_bn, uint8_t* factor_inverse, uint32_t* factor_fan_mask, uint32_t* pred_off, size_t pred_capacity, int32_t* pred_col,
               uint8_t* pred_kind, double* pred_a, double* pred_b, size_t* n_factors, size_t* n_preds) {
    unsigned n_thr = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("BC_SQLC_THREADS")) n_thr = (unsigned)std::atoi(e);
    if (n_thr < 1) n_thr = 1;
    if (n_thr > 64) n_thr = 64;
    if (n < 512) n_thr = 1;
    else if (n_thr > n / 256) n_thr = (unsigned)(n / 256);
    std::vector<Plan> part(n_thr);
    auto run = [&](auto&& fn) {
        if (n_thr == 1) { fn(0u); return; }
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_thr; ++t) pool.emplace_back(fn, t);
        for (std::thread& th : pool) th.join();
    };
    run([&](unsigned t) {
        Plan& p = part[t];
        const size_t q0 = n * t / n_thr, q1 = n * (t + 1) / n_thr;
        p.q_nfactors.reserve(q1 - q0);
        p.q_join.reserve(q1 - q0);
        p.q_status.reserve(q1 - q0);
        p.factors.reserve((q1 - q0) * 3);
        p.pcol.reserve((q1 - q0) * 10); p.pkind.reserve((q1 - q0) * 10); p.pa.reserve((q1 - q0) * 10); p.pb.reserve((q1 - q0) * 10);
        for (size_t q = q0; q < q1; ++q) {
            const size_t f0 = p.factors.size(), p0 = p.pcol.size();
            p.q_nfactors.push_back(0);
            p.q_join.push_back(0.0);
            const sv text = text_of(q);
            const bool ok = text.data() && plan_one(*h, text, p);
            if (!ok) {   // roll back what a half-planned query appended
                p.factors.resize(f0);
                p.pcol.resize(p0); p.pkind.resize(p0); p.pa.resize(p0); p.pb.resize(p0);
                p.q_nfactors.back() = 0;
            }
            p.q_status.push_back(ok ? 0 : 1);
        }
    });
    size_t nf = 0, np = 0;
    std::vector<size_t> f_base(n_thr), p_base(n_thr);
    for (unsigned t = 0; t < n_thr; ++t) {
        f_base[t] = nf;
        p_base[t] = np;
        nf += part[t].factors.size();
        np += part[t].pcol.size();
    }
    *n_factors = nf;
    *n_preds = np;
    if (nf > factor_capacity || np > pred_capacity || (nf && (!factor_bn || !factor_inverse || !factor_fan_mask || !pred_off)) ||
        (np && (!pred_col || !pred_kind || !pred_a || !pred_b))) {
        bc_set_error("bc_joblight_plan: %zu factors / %zu predicates do not fit the buffers", nf, np);
        return BC_ELIMIT;
    }
    if (nf > 0xFFFFFFFFull || np > 0xFFFFFFFFull) { bc_set_error("bc_joblight_plan: more than 2^32 factors / predicates in one batch"); return BC_ELIMIT; }
    run([&](unsigned t) {
        const Plan& p = part[t];
        size_t q = n * t / n_thr, fi = f_base[t], pi = p_base[t], lf = 0;
        for (size_t i = 0; i < p.q_status.size(); ++i, ++q) {
            status[q] = p.q_status[i];
            join_size[q] = p.q_join[i];
            first_factor[q] = (uint32_t)fi;
            for (uint32_t k = 0; k < p.q_nfactors[i]; ++k, ++lf, ++fi) {
                const Factor& fc = p.factors[lf];
                factor_bn[fi] = fc.bn;
                factor_inverse[fi] = fc.inverse;
                factor_fan_mask[fi] = fc.fan_mask;
                pred_off[fi] = (uint32_t)pi;
                pi += fc.n_pred;
            }
        }
        if (!p.pcol.empty()) {
            std::memcpy(pred_col + p_base[t], p.pcol.data(), p.pcol.size() * sizeof(int32_t));
            std::memcpy(pred_kind + p_base[t], p.pkind.data(), p.pkind.size());
            std::memcpy(pred_a + p_base[t], p.pa.data(), p.pa.size() * sizeof(double));
            std::memcpy(pred_b + p_base[t], p.pb.data(), p.pb.size() * sizeof(double));
        }
    });
    first_factor[n] = (uint32_t)nf;
    if (pred_off) pred_off[nf] = (uint32_t)np;
    return BC_OK;
}

}  // namespace

extern "C" {

int bc_joblight_plan(const bc_joblight* h, size_t n, const char* const* sqls, uint8_t* status, double* join_size, uint32_t* first_factor,
                     size_t factor_capacity, int32_t* factor_bn, uint8_t* factor_inverse, uint32_t* factor_fan_mask, uint32_t* pred_off,
                     size_t pred_capacity, int32_t* pred_col, uint8_t* pred_kind, double* pred_a, double* pred_b, size_t* n_factors,
                     size_t* n_preds) {
    if (!h || (n && (!sqls || !status || !join_size || !first_factor)) || !n_factors || !n_preds) { bc_set_error("bc_joblight_plan: bad arguments"); return BC_EINVAL; }
    return plan_batch(h, n, [&](size_t q) { return sqls[q] ? sv(sqls[q]) : sv(); }, status, join_size, first_factor, factor_capacity, factor_bn,
                      factor_inverse, factor_fan_mask, pred_off, pred_capacity, pred_col, pred_kind, pred_a, pred_b, n_factors, n_preds);
}

// The same over ONE text buffer: query q is text[text_off[q], text_off[q + 1]) (separators, if any, are whitespace to the parser).
// A batch of Python strings costs ~1 us each just to be turned into a char* array; one joined buffer does not.
int bc_joblight_plan_text(const bc_joblight* h, size_t n, const char* text, const uint64_t* text_off, uint8_t* status, double* join_size,
                          uint32_t* first_factor, size_t factor_capacity, int32_t* factor_bn, uint8_t* factor_inverse, uint32_t* factor_fan_mask,
                          uint32_t* pred_off, size_t pred_capacity, int32_t* pred_col, uint8_t* pred_kind, double* pred_a, double* pred_b,
                          size_t* n_factors, size_t* n_preds) {
    if (!h || (n && (!text || !text_off || !status || !join_size || !first_factor)) || !n_factors || !n_preds) { bc_set_error("bc_joblight_plan_text: bad arguments"); return BC_EINVAL; }
    for (size_t q = 0; q < n; ++q)
        if (text_off[q + 1] < text_off[q]) { bc_set_error("bc_joblight_plan_text: text_off is not monotone at query %zu", q); return BC_EINVAL; }
    return plan_batch(h, n, [&](size_t q) { return sv(text + text_off[q], (size_t)(text_off[q + 1] - text_off[q])); }, status, join_size, first_factor,
                      factor_capacity, factor_bn, factor_inverse, factor_fan_mask, pred_off, pred_capacity, pred_col, pred_kind, pred_a, pred_b,
                      n_factors, n_preds);
}

// BN_ensemble.cardinality (Models/BN_ensemble_model.py:228-252): card = join_size * prod(p | 1 / p); a factor of 0 makes the
// estimate 1; the result is clamped to >= 1.  status[q] != 0 (not planned here) leaves out[q] untouched.
int bc_joblight_combine(size_t n, const uint8_t* status, const double* join_size, const uint32_t* first_factor, const uint8_t* factor_inverse,
                        const double* factor_prob, double* out) {
    if (n && (!join_size || !first_factor || !out)) { bc_set_error("bc_joblight_combine: bad arguments"); return BC_EINVAL; }
    for (size_t q = 0; q < n; ++q) {
        if (status && status[q]) continue;
        double card = join_size[q];
        bool one = false;
        for (uint32_t f = first_factor[q]; f < first_factor[q + 1]; ++f) {
            const double p = factor_prob[f];
            if (p == 0) { one = true; break; }
            card = factor_inverse[f] ? card * (1 / p) : card * p;
        }
        out[q] = one || card <= 1 ? 1.0 : card;
    }
    return BC_OK;
}

}  // extern "C"
