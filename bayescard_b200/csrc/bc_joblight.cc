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
    std::map<std::pair<std::string, std::string>, double> rdc;   // pairwise RDC of "table.column" pairs (both orders)
    double epsilon = 0.1;
    std::map<std::string, std::string> alias;  // default alias -> table (the FROM clause overrides)
};

namespace {

struct Cond {
    int table;        // -1: title, else BN index
    std::string col;
    int op;           // 0 '=', 1 '<', 2 '>', 3 '<=', 4 '>='
    double val;
};

struct Factor {
    int32_t bn;
    uint8_t inverse;
    uint32_t fan_mask;
    uint32_t n_pred;
};

struct Plan {   // per thread
    std::vector<Factor> factors;
    std::vector<int32_t> pcol;
    std::vector<uint8_t> pkind;
    std::vector<double> pa, pb;
    std::vector<uint32_t> q_nfactors;   // per query
    std::vector<uint8_t> q_status;      // 0 ok, 1 not a job-light star query (left to the Python planner)
    std::vector<double> q_join;
};

sv strip(sv s) {
    while (!s.empty() && std::isspace((unsigned char)s.front())) s.remove_prefix(1);
    while (!s.empty() && std::isspace((unsigned char)s.back())) s.remove_suffix(1);
    return s;
}
bool ieq(sv a, const char* b) {
    const size_t n = std::strlen(b);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
// case-insensitive search of a keyword surrounded by whitespace
size_t find_kw(sv s, const char* kw, size_t from = 0) {
    const size_t n = std::strlen(kw);
    for (size_t i = from; i + n <= s.size(); ++i) {
        bool ok = true;
        for (size_t j = 0; j < n && ok; ++j) ok = std::tolower((unsigned char)s[i + j]) == std::tolower((unsigned char)kw[j]);
        if (ok && (i == 0 || std::isspace((unsigned char)s[i - 1])) && (i + n == s.size() || std::isspace((unsigned char)s[i + n]))) return i;
    }
    return sv::npos;
}
bool is_ident(sv s) {
    if (s.empty()) return false;
    for (char ch : s)
        if (!(std::isalnum((unsigned char)ch) || ch == '_')) return false;
    return true;
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
    std::map<std::string, int> alias_of;   // alias -> table id (-1 title, BN index otherwise)
    std::vector<int> order;                // joined tables (BN index) in FROM order
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
            while (sp < part.size() && !std::isspace((unsigned char)part[sp])) ++sp;
            const sv table = part.substr(0, sp);
            sv al = strip(part.substr(sp));
            if (al.size() > 3 && ieq(al.substr(0, 2), "as") && std::isspace((unsigned char)al[2])) al = strip(al.substr(3));   // "table AS alias"
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
            alias_of[std::string(al)] = id;
            if (id >= 0) order.push_back(id);
            if (comma == from.size()) break;
        }
    }
    if (!has_title || order.empty()) return false;
    // ---- WHERE: conditions "alias.col op number" joined by AND; join conditions "a.x = b.y" are skipped
    std::vector<Cond> conds;
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
            auto it = alias_of.find(std::string(lhs.substr(0, dot)));
            if (it == alias_of.end()) return false;
            // number: -?digits(.digits)?
            {
                size_t i = 0;
                if (i < rhs.size() && rhs[i] == '-') ++i;
                size_t d0 = i;
                while (i < rhs.size() && std::isdigit((unsigned char)rhs[i])) ++i;
                if (i == d0) return false;
                if (i < rhs.size() && rhs[i] == '.') {
                    ++i;
                    size_t d1 = i;
                    while (i < rhs.size() && std::isdigit((unsigned char)rhs[i])) ++i;
                    if (i == d1) return false;
                }
                if (i != rhs.size()) return false;
            }
            Cond cd;
            cd.table = it->second;
            cd.col = std::string(lhs.substr(dot + 1));
            cd.op = ops == "=" ? 0 : ops == "<" ? 1 : ops == ">" ? 2 : ops == "<=" ? 3 : ops == ">=" ? 4 : -1;
            if (cd.op < 0) return false;
            cd.val = std::strtod(std::string(rhs).c_str(), nullptr);
            conds.push_back(cd);
        }
    }
    // ---- per table: columns in first-mention order, each {eq | (lo, hi)}  (joblight.py _table_query)
    struct ColQ { std::string col; bool has_eq = false; double eq = 0, lo = -HUGE_VAL, hi = HUGE_VAL; };
    auto table_query = [&](int table) {
        std::vector<ColQ> q;
        for (const Cond& c : conds) {
            if (c.table != table) continue;
            ColQ* cq = nullptr;
            for (ColQ& x : q)
                if (x.col == c.col) cq = &x;
            if (!cq) { q.emplace_back(); cq = &q.back(); cq->col = c.col; }
            switch (c.op) {
                case 0: cq->has_eq = true; cq->eq = c.val; break;
                case 2: cq->lo = std::max(cq->lo, c.val + h.epsilon); break;
                case 4: cq->lo = std::max(cq->lo, c.val); break;
                case 1: cq->hi = std::min(cq->hi, c.val - h.epsilon); break;
                case 3: cq->hi = std::min(cq->hi, c.val); break;
            }
        }
        return q;
    };
    auto tname = [&](int table) -> const std::string& { static const std::string t = "title"; return table < 0 ? t : h.table[table]; };
    const std::vector<ColQ> c_t = table_query(-1);
    // ---- first model: _greedily_select_first_cardinality_spn with rdc_spn_selection (joblight.py _first_table)
    int first = -1;
    {
        std::vector<int> sorted = order;
        std::sort(sorted.begin(), sorted.end());
        sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
        double best_rdc = 0;
        int best_where = 0;
        for (int t : sorted) {
            std::vector<std::string> cols;   // set of conditioned "table.column" names over {title, t}
            auto add = [&](int tb) {
                for (const Cond& c : conds)
                    if (c.table == tb) {
                        std::string nm = tname(tb) + "." + c.col;
                        if (std::find(cols.begin(), cols.end(), nm) == cols.end()) cols.push_back(nm);
                    }
            };
            add(-1);
            add(t);
            double rdc = 0;
            for (const std::string& x : cols)
                for (const std::string& y : cols)
                    if (x < y) {
                        auto it = h.rdc.find({x, y});
                        if (it != h.rdc.end()) rdc += it->second;
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
    auto emit = [&](int bn, bool inverse, uint32_t fan_mask, const std::vector<const std::vector<ColQ>*>& parts, const std::vector<int>& part_table,
                    int nn_table) -> bool {
        Factor fc{bn, (uint8_t)inverse, fan_mask, 0};
        // dict semantics of the Python planner: a later update() of the same key replaces the value but keeps the position
        std::vector<std::string> names;
        std::vector<const ColQ*> vals;
        for (size_t i = 0; i < parts.size(); ++i)
            for (const ColQ& cq : *parts[i]) {
                std::string nm = tname(part_table[i]) + "." + cq.col;
                size_t j = 0;
                for (; j < names.size(); ++j)
                    if (names[j] == nm) break;
                if (j == names.size()) { names.push_back(nm); vals.push_back(&cq); }
                else vals[j] = &cq;
            }
        for (size_t j = 0; j < names.size(); ++j) {
            const int ci = bc_sqlc_column_index(h.sqlc[bn], names[j].c_str());
            out.pcol.push_back(ci);   // -1: KeyError in the mirror -> the factor compiler flags the factor
            out.pkind.push_back(vals[j]->has_eq ? 0 : 1);
            out.pa.push_back(vals[j]->has_eq ? vals[j]->eq : vals[j]->lo);
            out.pb.push_back(vals[j]->has_eq ? 0.0 : vals[j]->hi);
            ++fc.n_pred;
        }
        {   // the NOT NULL condition of relevant_conditions: <table>.<table>_nn = 1
            const std::string nm = h.table[nn_table] + "." + h.table[nn_table] + "_nn";
            out.pcol.push_back(bc_sqlc_column_index(h.sqlc[bn], nm.c_str()));
            out.pkind.push_back(0);
            out.pa.push_back(1.0);
            out.pb.push_back(0.0);
            ++fc.n_pred;
        }
        out.factors.push_back(fc);
        return true;
    };
    uint32_t nf = 0;
    {
        const std::vector<ColQ> c_a = table_query(first);
        uint32_t mask = 0;
        for (int b : order)
            if (b != first) {
                const int node = h.fan_node[(size_t)first * h.n_bn + b];
                if (node < 0 || node >= 32) return false;   // no such fan-out column in the first model: leave it to Python
                mask |= 1u << node;
            }
        emit(first, false, mask, {&c_t, &c_a}, {-1, first}, first);
        ++nf;
    }
    std::vector<int> seen;
    for (int b : order) {
        if (b == first || std::find(seen.begin(), seen.end(), b) != seen.end()) continue;
        seen.push_back(b);
        const std::vector<ColQ> c_b = table_query(b);
        if (c_b.empty()) continue;   // factor_refine: nominator and denominator cancel
        emit(b, false, 0, {&c_t, &c_b}, {-1, b}, b);
        emit(b, true, 0, {&c_t}, {-1}, b);
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
    for (int i = 0; i < n_rdc; ++i) {
        h->rdc[{rdc_a[i], rdc_b[i]}] = rdc_val[i];
        h->rdc[{rdc_b[i], rdc_a[i]}] = rdc_val[i];
    }
    h->epsilon = epsilon;
    *out = h;
    return BC_OK;
}

void bc_joblight_destroy(bc_joblight* h) { delete h; }

int bc_joblight_plan(const bc_joblight* h, size_t n, const char* const* sqls, uint8_t* status, double* join_size, uint32_t* first_factor,
                     size_t factor_capacity, int32_t* factor_bn, uint8_t* factor_inverse, uint32_t* factor_fan_mask, uint32_t* pred_off,
                     size_t pred_capacity, int32_t* pred_col, uint8_t* pred_kind, double* pred_a, double* pred_b, size_t* n_factors,
                     size_t* n_preds) {
    if (!h || (n && (!sqls || !status || !join_size || !first_factor)) || !n_factors || !n_preds) { bc_set_error("bc_joblight_plan: bad arguments"); return BC_EINVAL; }
    unsigned n_thr = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("BC_SQLC_THREADS")) n_thr = (unsigned)std::atoi(e);
    if (n_thr < 1) n_thr = 1;
    if (n_thr > 64) n_thr = 64;
    if (n < 512) n_thr = 1;
    else if (n_thr > n / 256) n_thr = (unsigned)(n / 256);
    std::vector<Plan> part(n_thr);
    auto work = [&](unsigned t) {
        Plan& p = part[t];
        const size_t q0 = n * t / n_thr, q1 = n * (t + 1) / n_thr;
        for (size_t q = q0; q < q1; ++q) {
            const size_t f0 = p.factors.size(), p0 = p.pcol.size();
            p.q_nfactors.push_back(0);
            p.q_join.push_back(0.0);
            const bool ok = sqls[q] && plan_one(*h, sv(sqls[q]), p);
            if (!ok) {   // roll back what a half-planned query appended
                p.factors.resize(f0);
                p.pcol.resize(p0); p.pkind.resize(p0); p.pa.resize(p0); p.pb.resize(p0);
                p.q_nfactors.back() = 0;
            }
            p.q_status.push_back(ok ? 0 : 1);
        }
    };
    if (n_thr == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_thr; ++t) pool.emplace_back(work, t);
        for (std::thread& th : pool) th.join();
    }
    size_t nf = 0, np = 0;
    for (const Plan& p : part) { nf += p.factors.size(); np += p.pcol.size(); }
    *n_factors = nf;
    *n_preds = np;
    if (nf > factor_capacity || np > pred_capacity || (nf && (!factor_bn || !factor_inverse || !factor_fan_mask || !pred_off)) ||
        (np && (!pred_col || !pred_kind || !pred_a || !pred_b))) {
        bc_set_error("bc_joblight_plan: %zu factors / %zu predicates do not fit the buffers", nf, np);
        return BC_ELIMIT;
    }
    size_t q = 0, fi = 0, pi = 0;
    for (const Plan& p : part) {
        size_t lf = 0, lp = 0;
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
                for (uint32_t j = 0; j < fc.n_pred; ++j, ++lp, ++pi) {
                    pred_col[pi] = p.pcol[lp];
                    pred_kind[pi] = p.pkind[lp];
                    pred_a[pi] = p.pa[lp];
                    pred_b[pi] = p.pb[lp];
                }
            }
        }
    }
    first_factor[n] = (uint32_t)fi;
    if (pred_off) pred_off[fi] = (uint32_t)pi;
    return BC_OK;
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
