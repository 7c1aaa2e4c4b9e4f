// Batched SQL -> descriptor compiler (host only): the predicate compiler of north_star (2), off the Python loop.
//
// One call turns N single-table SQL texts ("SELECT COUNT(*) FROM t WHERE a OP v AND ...") into the per-query
// descriptor rows the CUDA kernels read.  It restates, for the common predicate shapes, the two reference
// functions that sit in front of the hot path:
//
//   parse_query_single_table   Evaluation/cardinality_estimation.py:22-119
//       split on " WHERE " / " AND " (:113-118); "col IN [..]" (:26-36) or the first run of [<>=] as operator,
//       operand tried as int, float, bare word (:38-57); continuous columns collect a (lo, hi) interval, strict
//       bounds moved by 1e-6 (:63-91); categorical columns collect the ORIGINAL values that satisfy the predicate,
//       inequalities being evaluated against BN.domain[attr] (:93-103); repeated predicates intersect (:105-109)
//   Bayescard_BN.query_decoding   Models/Bayescard_BN.py:279-325 with realign (:53-72), continuous_range_map
//       (:180-239, including its one-sided-stop quirk) and apply_encoding_to_value / apply_ndistinct_to_value
//       (Models/BN_single_model.py:98-139): values -> (bin, weight), duplicate bins add up capped at 1.
//
// Everything unusual -- predicates the reference would raise on, operands whose Python parsing is subtle
// (underscores, inf / nan, leading zeros, quoted lists with escapes) -- is NOT guessed at: the query is flagged
// BC_SQLC_PYTHON and the host library (bayescard_b200/sqlc.py) runs it through its Python mirror of the same two
// functions, so results are identical by construction.  tests/test_sqlc.py compares every row this file emits with
// that mirror on the shipped workloads and on fuzzed SQL.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "bc_internal.h"

namespace {

// A value of a predicate or of a domain.  `str` points into the SQL text being compiled or into the column's own
// tables; both outlive the query.
struct Val {
    bool is_str = false;
    double num = 0.0;
    std::string_view str;
    bool operator==(const Val& o) const { return is_str == o.is_str && (is_str ? str == o.str : num == o.num); }
};

uint64_t num_key(double x) {
    if (x == 0.0) x = 0.0;  // -0.0 and 0.0 are the same dict key
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return b;
}

struct Column {
    std::string name;
    int node = -1;           // topological index, -1: known to the table but outside the tree
    bool continuous = false;
    int card = 0;
    // categorical
    bool has_encoding = false;
    std::vector<int32_t> enc_bin;
    std::vector<double> enc_w;
    std::unordered_map<uint64_t, int> enc_num;       // numeric original value -> entry
    std::deque<std::string> text;                    // owns the strings the views below point into (stable addresses)
    std::unordered_map<std::string_view, int> enc_str;  // string original value -> entry
    std::vector<Val> domain;                         // BN.domain[attr] in its stored order
    bool domain_numeric = false;                     // every domain value is a number (inequalities are defined)
    std::vector<double> enc_key;                     // the encoding's keys in dict order (meaningful when enc_all_num)
    bool enc_all_num = true;                         // every key of the encoding is a number: (lo, hi) tuples are defined
    bool has_null = false;                           // BN.null_values[attr]: skipped by (lo, hi) tuples (Bayescard_BN.py:304-318)
    double null_num = 0;
    // continuous
    double dom_lo = 0, dom_hi = 0;
    std::vector<double> edge_lo, edge_hi;
    std::unordered_map<uint64_t, double> nd_map;
};

// state of one column inside one query, while its predicates are folded
struct ColState {
    int col = -1;
    // categorical: the current value list
    std::vector<Val> vals;
    // continuous: interval or point
    bool is_point = false;
    double lo = 0, hi = 0, point = 0;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

using sv = std::string_view;

sv strip(sv s) {
    size_t a = 0, b = s.size();
    while (a < b && is_space(s[a])) ++a;
    while (b > a && is_space(s[b - 1])) --b;
    return s.substr(a, b - a);
}

void split_on(sv s, sv sep, std::vector<sv>& out) {
    out.clear();
    size_t pos = 0;
    for (;;) {
        const size_t f = s.find(sep, pos);
        if (f == sv::npos) {
            out.push_back(s.substr(pos));
            return;
        }
        out.push_back(s.substr(pos, f - pos));
        pos = f + sep.size();
    }
}

bool ieq(sv a, const char* b) {
    size_t i = 0;
    for (; i < a.size() && b[i]; ++i)
        if (std::tolower((unsigned char)a[i]) != b[i]) return false;
    return i == a.size() && !b[i];
}

// Classify an operand the way Python's int() / float() would, conservatively.
//   0 = number (value in *out), 1 = plain string, 2 = let Python decide
int classify_scalar(sv t, double* out) {
    if (t.empty()) return 1;
    size_t i = 0;
    if (t[i] == '+' || t[i] == '-') ++i;
    if (i == t.size()) return 1;
    bool digits = false, dot = false, expo = false, other = false;
    size_t int_digits = 0;
    for (size_t j = i; j < t.size(); ++j) {
        const char c = t[j];
        if (c >= '0' && c <= '9') {
            digits = true;
            if (!dot && !expo) ++int_digits;
        } else if (c == '.' && !dot && !expo) {
            dot = true;
        } else if ((c == 'e' || c == 'E') && digits && !expo && j + 1 < t.size()) {
            expo = true;
            if (t[j + 1] == '+' || t[j + 1] == '-') ++j;
            if (j + 1 >= t.size()) other = true;
        } else {
            other = true;
            break;
        }
    }
    if (!other && digits) {
        if (!dot && !expo && int_digits > 15) return 2;                         // beyond exact doubles
        if (!dot && !expo && int_digits > 1 && t[i] == '0') return 2;           // "007": int() takes it, literals do not
        char buf[80];
        if (t.size() >= sizeof(buf)) return 2;
        std::memcpy(buf, t.data(), t.size());
        buf[t.size()] = 0;
        char* end = nullptr;
        *out = std::strtod(buf, &end);
        if (end != buf + t.size()) return 2;
        return 0;
    }
    // words float() accepts, digits with underscores, hex ...: not worth restating
    const sv w = t.substr(i);
    if (ieq(w, "inf") || ieq(w, "infinity") || ieq(w, "nan")) return 2;
    if (digits && t.find('_') != sv::npos) return 2;
    return 1;
}

}  // namespace

struct bc_sqlc {
    int n_nodes = 0;
    std::vector<int32_t> card;
    std::vector<int64_t> bit_off, dense_off;
    int64_t bits_row_bytes = 0, dense_width = 0, total_bits = 0;
    std::vector<uint8_t> bits_default;   // every state of every column selected
    std::vector<float> dense_default;    // weight 1 on every state, 0 on the padding
    std::deque<Column> cols;                        // built in place: the views inside a Column never move
    std::deque<std::string> names;                  // owns the keys of by_name
    std::unordered_map<std::string_view, int> by_name;
};

namespace {

enum { KIND_BITS = 0, KIND_DENSE = 1, KIND_ZERO = 2, KIND_PYTHON = 3, KIND_OVERFLOW = 4 };

struct Decoded {
    int node;
    std::vector<int32_t> bins;
    std::vector<double> wts;
};

// continuous_range_map (Models/Bayescard_BN.py:180-239): bins overlapped by [lo, hi] with fractional coverage
void continuous_bins(const Column& c, double lo, double hi, std::vector<int32_t>& bins, std::vector<double>& cov) {
    const int n = (int)c.edge_lo.size();
    auto cover = [&](int k) -> double {
        const double tl = c.edge_lo[k], tr = c.edge_hi[k];
        if (lo >= tr || hi <= tl) return 0.0;
        if (hi > tr) return lo < tl ? 1.0 : (tr - lo) / (tr - tl);
        return lo > tl ? (hi - lo) / (tr - tl) : (hi - tl) / (tr - tl);
    };
    int i = 0, j = n;
    while (i != j) {
        const int mid = (int)(i + (double)(j - i) / 2);
        const double tl = c.edge_lo[mid], tr = c.edge_hi[mid];
        if (lo >= tr) {
            if (i == mid) break;
            i = mid;
        } else if (hi <= tl) {
            j = mid;
        } else {
            i = j = mid;
        }
    }
    int down = i, up = i + 1;
    bool more_down = true, more_up = true;
    bins.clear();
    cov.clear();
    while (down >= 0 && up < n && (more_down || more_up)) {
        if (more_down) {
            const double cv = cover(down);
            if (cv != 0) { bins.push_back(down); cov.push_back(cv); --down; }
            else more_down = false;
        }
        if (more_up) {
            const double cv = cover(up);
            if (cv != 0) { bins.push_back(up); cov.push_back(cv); ++up; }
            else more_up = false;
        }
    }
}

// Parses "[a, b, ...]".  0 = ok, 2 = let Python decide
int parse_in_list(sv body, std::vector<Val>& out, std::vector<sv>& items) {
    out.clear();
    if (body.size() < 2 || body.front() != '[' || body.back() != ']') return 2;
    const sv inner = body.substr(1, body.size() - 2);
    for (char ch : inner)
        if (ch == '[' || ch == ']' || ch == '(' || ch == ')' || ch == '{' || ch == '}' || ch == '\\' || ch == '#') return 2;
    if (strip(inner).empty()) return 0;  // [] : an empty value list (undecodable -> estimate 0)
    const bool quoted = inner.find('\'') != sv::npos || inner.find('"') != sv::npos;
    if (!quoted) {
        split_on(inner, ",", items);
        // ast.literal_eval succeeds only when every item is a literal (one trailing comma allowed); one bare word
        // or a stray empty item makes it raise, and then the reference splits on commas and keeps the raw text of
        // EVERY item, empty ones included ('' is a real value of DMV's Fuel_Type)
        bool all_num = true, any_text = false;
        std::vector<double> nums;
        for (size_t i = 0; i < items.size(); ++i) {
            sv& it = items[i];
            it = strip(it);
            if (it == "True" || it == "False" || it == "None") return 2;
            if (it.empty()) {
                if (i + 1 != items.size() || items.size() == 1) all_num = false;  // not a lone trailing comma
                continue;
            }
            any_text = true;
            double x = 0;
            const int k = classify_scalar(it, &x);
            if (k == 2) return 2;
            if (k == 0) nums.push_back(x);
            else all_num = false;
        }
        if (!any_text) return 2;  // "[,]" and friends
        if (all_num) {
            for (double x : nums) { Val v; v.num = x; out.push_back(v); }
        } else {
            for (const sv it : items) { Val v; v.is_str = true; v.str = it; out.push_back(v); }
        }
        return 0;
    }
    // quoted items: accept only  'text' / "text" / number , separated by commas outside the quotes
    size_t p = 0;
    const size_t n = inner.size();
    for (;;) {
        while (p < n && is_space(inner[p])) ++p;
        if (p >= n) return 2;  // trailing comma
        Val v;
        if (inner[p] == '\'' || inner[p] == '"') {
            const char q = inner[p];
            const size_t e = inner.find(q, p + 1);
            if (e == sv::npos) return 2;
            v.is_str = true;
            v.str = inner.substr(p + 1, e - p - 1);
            p = e + 1;
        } else {
            size_t e = inner.find(',', p);
            if (e == sv::npos) e = n;
            const sv it = strip(inner.substr(p, e - p));
            double x = 0;
            if (classify_scalar(it, &x) != 0) return 2;  // a bare word next to quoted strings: literal_eval fails
            v.num = x;
            p = e;
        }
        out.push_back(v);
        while (p < n && is_space(inner[p])) ++p;
        if (p >= n) return 0;
        if (inner[p] != ',') return 2;
        ++p;
    }
}

int cmp_op(sv op) {
    if (op == ">") return 0;
    if (op == "<") return 1;
    if (op == ">=") return 2;
    if (op == "<=") return 3;
    if (op == "=" || op == "==") return 4;
    return -1;
}

// Scratch reused from query to query (the inner vectors keep their capacity: no allocation in steady state).
struct Scratch {
    std::vector<Decoded> dec;
    size_t n_dec = 0;
    std::vector<ColState> state;
    size_t n_state = 0;
    std::vector<sv> preds, parts, items;
    std::vector<Val> list, dom, both;
    std::vector<double> seg;
};

// One query.  Returns the kind; fills sc.dec[0 .. n_dec) (columns inside the tree only) for KIND_BITS / KIND_DENSE.
int compile_one(const bc_sqlc& c, sv sql, Scratch& sc) {
    const double kInf = HUGE_VAL;
    split_on(sql, " WHERE ", sc.parts);
    if (sc.parts.size() < 2) return KIND_PYTHON;
    const sv where = strip(sc.parts.back());
    split_on(where, " AND ", sc.preds);
    sc.n_state = 0;
    auto find_state = [&](int col) -> ColState* {
        for (size_t i = 0; i < sc.n_state; ++i)
            if (sc.state[i].col == col) return &sc.state[i];
        return nullptr;
    };
    auto new_state = [&](int col) -> ColState& {
        if (sc.n_state == sc.state.size()) sc.state.emplace_back();
        ColState& s = sc.state[sc.n_state++];
        s.col = col;
        s.vals.clear();
        s.is_point = false;
        return s;
    };
    for (const sv raw : sc.preds) {
        const sv pred = strip(raw);
        sv attr, op;
        Val scalar;
        bool is_in = false;
        split_on(pred, " IN ", sc.parts);
        if (sc.parts.size() > 2) return KIND_PYTHON;
        if (sc.parts.size() == 2) {
            attr = strip(sc.parts[0]);
            is_in = true;
            if (parse_in_list(strip(sc.parts[1]), sc.list, sc.items) != 0) return KIND_PYTHON;
        } else {
            size_t first = sv::npos;
            for (size_t i = 0; i < pred.size(); ++i)
                if (pred[i] == '<' || pred[i] == '>' || pred[i] == '=') { first = i; break; }
            if (first == sv::npos) return KIND_PYTHON;  // NameError in the reference
            const size_t last = (first + 1 < pred.size() && (pred[first + 1] == '<' || pred[first + 1] == '>' || pred[first + 1] == '=')) ? first + 1 : first;
            attr = strip(pred.substr(0, first));
            op = pred.substr(first, last - first + 1);
            const sv text = strip(pred.substr(last + 1));
            double x = 0;
            const int k = classify_scalar(text, &x);
            if (k == 2) return KIND_PYTHON;
            if (k == 0) scalar.num = x;
            else { scalar.is_str = true; scalar.str = text; }
        }
        auto it = c.by_name.find(attr);
        if (it == c.by_name.end()) continue;  // not a column of this BN: the predicate is dropped (:59-60)
        const Column& col = c.cols[it->second];
        ColState* st = find_state(it->second);
        if (col.continuous) {
            if (is_in || scalar.is_str) return KIND_PYTHON;  // assertion / TypeError in the reference
            const int o = cmp_op(op);
            if (o < 0) return KIND_PYTHON;
            double lo = -kInf, hi = kInf;
            if (o == 2) lo = scalar.num;
            else if (o == 0) lo = scalar.num + 1e-6;
            else if (o == 3) hi = scalar.num;
            else if (o == 1) hi = scalar.num - 1e-6;
            if (st) {
                // a point predicate next to another one: the reference indexes a float (TypeError)
                if (st->is_point || o == 4) return KIND_PYTHON;
                st->lo = std::max(st->lo, lo);
                st->hi = std::min(st->hi, hi);
                continue;
            }
            ColState& ns = new_state(it->second);
            if (o == 4) { ns.is_point = true; ns.point = scalar.num; }
            else { ns.lo = lo; ns.hi = hi; }
            continue;
        }
        // categorical (and boolean) columns
        if (col.domain.empty()) return KIND_PYTHON;  // the reference reads domain[0] before it looks at the operator
        std::vector<Val>& dom = is_in ? sc.list : sc.dom;
        if (!is_in) {
            dom.clear();
            const int o = cmp_op(op);
            if (o < 0) return KIND_PYTHON;
            if (o == 4) {
                dom.push_back(scalar);
            } else {
                if (!col.domain_numeric || scalar.is_str) return KIND_PYTHON;
                for (const Val& d : col.domain) {
                    const bool keep = o == 0 ? d.num > scalar.num : o == 1 ? d.num < scalar.num : o == 2 ? d.num >= scalar.num : d.num <= scalar.num;
                    if (keep) dom.push_back(d);
                }
            }
        }
        if (st) {  // repeated predicate: keep the values that the earlier list holds too (:105-109)
            sc.both.clear();
            for (const Val& v : dom)
                for (const Val& e : st->vals)
                    if (v == e) { sc.both.push_back(v); break; }
            st->vals.assign(sc.both.begin(), sc.both.end());
        } else {
            new_state(it->second).vals.assign(dom.begin(), dom.end());
        }
    }
    // ---- query_decoding
    sc.n_dec = 0;
    int in_tree = 0;
    Decoded tmp;
    for (size_t si = 0; si < sc.n_state; ++si) {
        const ColState& s = sc.state[si];
        const Column& col = c.cols[s.col];
        Decoded* dp = &tmp;
        if (col.node >= 0) {
            if (sc.n_dec == sc.dec.size()) sc.dec.emplace_back();
            dp = &sc.dec[sc.n_dec];
        }
        Decoded& d = *dp;
        d.node = col.node;
        d.bins.clear();
        d.wts.clear();
        if (col.continuous) {
            double lo, hi, mult = 1.0;
            bool has_mult = false;
            if (s.is_point) {
                lo = s.point - 0.5;
                hi = s.point + 0.5;
                auto f = col.nd_map.find(num_key(s.point));
                if (f != col.nd_map.end()) { mult = f->second; has_mult = true; }
            } else {
                lo = std::max(col.dom_lo, s.lo);
                hi = std::min(col.dom_hi, s.hi);
            }
            if (lo > hi) return KIND_ZERO;
            if (col.edge_lo.empty()) return KIND_PYTHON;
            continuous_bins(col, lo, hi, d.bins, d.wts);
            if (has_mult)
                for (double& w : d.wts) w *= mult;
        } else {
            if (!col.has_encoding) return KIND_ZERO;
            if (s.vals.empty()) return KIND_ZERO;
            // realign: drop unknown values, first-occurrence order, duplicate bins add up capped at 1
            for (const Val& v : s.vals) {
                int e = -1;
                if (v.is_str) {
                    auto f = col.enc_str.find(v.str);
                    if (f != col.enc_str.end()) e = f->second;
                } else {
                    if (std::isnan(v.num)) continue;
                    auto f = col.enc_num.find(num_key(v.num));
                    if (f != col.enc_num.end()) e = f->second;
                }
                if (e < 0) continue;
                const int32_t b = col.enc_bin[e];
                const double w = col.enc_w[e];
                size_t j = 0;
                for (; j < d.bins.size(); ++j)
                    if (d.bins[j] == b) break;
                if (j == d.bins.size()) { d.bins.push_back(b); d.wts.push_back(w); }
                else d.wts[j] = std::min(d.wts[j] + w, 1.0);
            }
        }
        if (col.node >= 0) {
            ++in_tree;
            for (int32_t b : d.bins)
                if (b < 0 || b >= col.card) return KIND_PYTHON;
            ++sc.n_dec;
        }
    }
    if (in_tree == 0) return KIND_ZERO;  // no queried column is reachable: ExactInference.py:197
    for (size_t i = 0; i < sc.n_dec; ++i)
        for (double w : sc.dec[i].wts)
            if (w != 1.0) return KIND_DENSE;
    return KIND_BITS;
}

// One FACTOR of a join query (Models/BN_ensemble_model.py:192-225 hands them to query_decoding as dicts
// {column: scalar | (lo, hi)}): predicates pred[p0 .. p1) = (column index of this compiler, kind 0 scalar / 1 tuple, a, b),
// numeric values only, one predicate per column.  Restates Bayescard_BN.query_decoding (Models/Bayescard_BN.py:279-325) for
// these two value shapes; fills sc.dec like compile_one.  `has_fan`: the factor carries fan-out columns (an expectation),
// which keeps it alive when no predicated column is reachable.
int compile_factor(const bc_sqlc& c, const int32_t* pcol, const uint8_t* pkind, const double* pa, const double* pb, uint32_t p0, uint32_t p1,
                   bool has_fan, Scratch& sc) {
    sc.n_dec = 0;
    int in_tree = 0;
    Decoded tmp;
    for (uint32_t p = p0; p < p1; ++p) {
        if (pcol[p] < 0 || pcol[p] >= (int32_t)c.cols.size()) return KIND_PYTHON;   // KeyError in the reference
        const Column& col = c.cols[pcol[p]];
        Decoded* dp = &tmp;
        if (col.node >= 0) {
            if (sc.n_dec == sc.dec.size()) sc.dec.emplace_back();
            dp = &sc.dec[sc.n_dec];
        }
        Decoded& d = *dp;
        d.node = col.node;
        d.bins.clear();
        d.wts.clear();
        const double a = pa[p], b = pb[p];
        if (col.continuous) {
            double lo, hi, mult = 1.0;
            bool has_mult = false;
            if (pkind[p] == 1) {
                lo = std::max(col.dom_lo, a);
                hi = std::min(col.dom_hi, b);
            } else {
                lo = a - 0.5;
                hi = a + 0.5;
                auto f = col.nd_map.find(num_key(a));
                if (f != col.nd_map.end()) { mult = f->second; has_mult = true; }
            }
            if (lo > hi) return KIND_ZERO;
            if (col.edge_lo.empty()) return KIND_PYTHON;
            continuous_bins(col, lo, hi, d.bins, d.wts);
            if (has_mult)
                for (double& w : d.wts) w *= mult;
        } else if (pkind[p] == 1) {
            // categorical (lo, hi): the ORIGINAL values of the encoding inside the range, in dict order, the null marker excluded
            if (!col.enc_all_num) return KIND_PYTHON;
            for (size_t e = 0; e < col.enc_key.size(); ++e) {
                const double k = col.enc_key[e];
                if (!(k >= a && k <= b)) continue;
                if (col.has_null && k == col.null_num) continue;
                const int32_t bin = col.enc_bin[e];
                const double w = col.enc_w[e];
                size_t j = 0;
                for (; j < d.bins.size(); ++j)
                    if (d.bins[j] == bin) break;
                if (j == d.bins.size()) { d.bins.push_back(bin); d.wts.push_back(w); }
                else d.wts[j] = std::min(d.wts[j] + w, 1.0);
            }
            if (d.bins.empty()) return KIND_ZERO;
        } else {
            if (!col.has_encoding) return KIND_ZERO;
            auto f = col.enc_num.find(num_key(a));
            if (f == col.enc_num.end()) return KIND_ZERO;
            d.bins.push_back(col.enc_bin[f->second]);
            d.wts.push_back(col.enc_w[f->second]);
        }
        if (col.node >= 0) {
            ++in_tree;
            for (int32_t bin : d.bins)
                if (bin < 0 || bin >= col.card) return KIND_PYTHON;
            ++sc.n_dec;
        }
    }
    if (in_tree == 0 && !has_fan) return KIND_ZERO;  // no queried column is reachable: ExactInference.py:197
    for (size_t i = 0; i < sc.n_dec; ++i)
        for (double w : sc.dec[i].wts)
            if (w != 1.0) return KIND_DENSE;
    return KIND_BITS;
}

// clear bits [o, o + len) of a little-endian bit row
inline void clear_bits(uint8_t* row, int64_t o, int len) {
    int64_t b = o, e = o + len;
    while (b < e && (b & 7)) { row[b >> 3] &= (uint8_t)~(1u << (b & 7)); ++b; }
    while (b + 8 <= e) { row[b >> 3] = 0; b += 8; }
    while (b < e) { row[b >> 3] &= (uint8_t)~(1u << (b & 7)); ++b; }
}

}  // namespace

extern "C" {

int bc_sqlc_create(int n_nodes, const int32_t* card, bc_sqlc** out) {
    if (!out || n_nodes <= 0 || !card) {
        bc_set_error("bc_sqlc_create: bad arguments");
        return BC_EINVAL;
    }
    bc_sqlc* c = new bc_sqlc();
    c->n_nodes = n_nodes;
    c->card.assign(card, card + n_nodes);
    int64_t bits = 0, dense = 0;
    for (int v = 0; v < n_nodes; ++v) {
        if (card[v] <= 0) {
            delete c;
            bc_set_error("bc_sqlc_create: node %d has %d states", v, card[v]);
            return BC_EINVAL;
        }
        c->bit_off.push_back(bits);
        c->dense_off.push_back(dense);
        bits += card[v];
        dense += bc_round_up(card[v], 4);
    }
    c->total_bits = bits;
    c->bits_row_bytes = bc_round_up(bits, 128) / 8;
    c->dense_width = dense;
    c->bits_default.assign((size_t)c->bits_row_bytes, 0);
    for (int64_t b = 0; b < bits; ++b) c->bits_default[b >> 3] |= (uint8_t)(1u << (b & 7));
    c->dense_default.assign((size_t)dense, 0.f);
    for (int v = 0; v < n_nodes; ++v)
        for (int s = 0; s < card[v]; ++s) c->dense_default[c->dense_off[v] + s] = 1.f;
    *out = c;
    return BC_OK;
}

void bc_sqlc_destroy(bc_sqlc* c) { delete c; }

int64_t bc_sqlc_bits_stride(const bc_sqlc* c) { return c ? c->bits_row_bytes : 0; }
int64_t bc_sqlc_dense_width(const bc_sqlc* c) { return c ? c->dense_width : 0; }

int bc_sqlc_add_categorical(bc_sqlc* c, const char* name, int node, int has_encoding, int n_enc, const uint8_t* enc_is_str,
                            const double* enc_num, const char* const* enc_str, const int32_t* enc_bin, const double* enc_weight,
                            int n_dom, const uint8_t* dom_is_str, const double* dom_num, const char* const* dom_str) {
    if (!c || !name || node >= c->n_nodes || n_enc < 0 || n_dom < 0) {
        bc_set_error("bc_sqlc_add_categorical: bad arguments");
        return BC_EINVAL;
    }
    c->cols.emplace_back();
    Column& col = c->cols.back();
    col.name = name;
    col.node = node;
    col.card = node >= 0 ? c->card[node] : 0;
    col.has_encoding = has_encoding != 0;
    for (int i = 0; i < n_enc; ++i) {
        col.enc_bin.push_back(enc_bin[i]);
        col.enc_w.push_back(enc_weight[i]);
        col.enc_key.push_back(enc_is_str[i] ? 0.0 : enc_num[i]);
        if (enc_is_str[i]) col.enc_all_num = false;
        if (enc_is_str[i]) {
            col.text.emplace_back(enc_str[i]);
            col.enc_str.emplace(std::string_view(col.text.back()), i);
        } else {
            col.enc_num.emplace(num_key(enc_num[i]), i);
        }
    }
    col.domain_numeric = n_dom > 0;
    for (int i = 0; i < n_dom; ++i) {
        Val v;
        v.is_str = dom_is_str[i] != 0;
        if (v.is_str) {
            col.text.emplace_back(dom_str[i]);
            v.str = col.text.back();
            col.domain_numeric = false;
        } else {
            v.num = dom_num[i];
        }
        col.domain.push_back(v);
    }
    c->names.push_back(col.name);
    c->by_name[std::string_view(c->names.back())] = (int)c->cols.size() - 1;
    return BC_OK;
}

int bc_sqlc_add_continuous(bc_sqlc* c, const char* name, int node, double dom_lo, double dom_hi, int n_bins,
                           const double* edge_lo, const double* edge_hi, int n_ndmap, const double* nd_key,
                           const double* nd_mult) {
    if (!c || !name || node >= c->n_nodes || n_bins < 0 || n_ndmap < 0) {
        bc_set_error("bc_sqlc_add_continuous: bad arguments");
        return BC_EINVAL;
    }
    c->cols.emplace_back();
    Column& col = c->cols.back();
    col.name = name;
    col.node = node;
    col.card = node >= 0 ? c->card[node] : 0;
    col.continuous = true;
    col.dom_lo = dom_lo;
    col.dom_hi = dom_hi;
    col.edge_lo.assign(edge_lo, edge_lo + n_bins);
    col.edge_hi.assign(edge_hi, edge_hi + n_bins);
    for (int i = 0; i < n_ndmap; ++i) col.nd_map.emplace(num_key(nd_key[i]), nd_mult[i]);
    c->names.push_back(col.name);
    c->by_name[std::string_view(c->names.back())] = (int)c->cols.size() - 1;
    return BC_OK;
}

}  // extern "C"

namespace {
// Batch driver shared by the SQL and the factor entry points: `one(q, scratch)` decodes query q into scratch.dec and returns
// its kind.  Queries are independent: slices of the batch are compiled on host threads.  BITS rows land at their query
// index; DENSE rows are collected per thread and appended in query order afterwards.
struct WsOut {   // DENSE-kind queries as WSPARSE rows (weighted runs, include/bayescard_b200.h) instead of DENSE_F32 rows
    uint32_t* row_off;   // [dense_capacity + 1]
    uint32_t* words;
    size_t capacity;     // words
    size_t* n_words;
};

template <class One>
int run_batch(const bc_sqlc* c, size_t n_queries, One one, uint8_t* kind, void* bits_rows, float* dense_rows, size_t dense_capacity,
              uint32_t* dense_index, size_t* n_dense, const WsOut* ws = nullptr) {
    uint8_t* bits = static_cast<uint8_t*>(bits_rows);
    unsigned n_thr = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("BC_SQLC_THREADS")) n_thr = (unsigned)std::atoi(e);
    if (n_thr < 1) n_thr = 1;
    if (n_thr > 64) n_thr = 64;
    if (n_queries < 512) n_thr = 1;
    else if (n_thr > n_queries / 256) n_thr = (unsigned)(n_queries / 256);
    struct Part {
        std::vector<float> dense;       // DENSE rows of this slice, in query order
        std::vector<uint32_t> index;
        std::vector<uint32_t> ws, ws_len;   // ... or their WSPARSE rows and row lengths
    };
    std::vector<Part> part(n_thr);
    const size_t width = (size_t)c->dense_width;
    auto work = [&](unsigned t) {
        Scratch sc;
        Part& pt = part[t];
        const size_t q0 = n_queries * t / n_thr, q1 = n_queries * (t + 1) / n_thr;
        for (size_t q = q0; q < q1; ++q) {
            const int k = one(q, sc);
            if (k == KIND_BITS) {
                uint8_t* row = bits + q * (size_t)c->bits_row_bytes;
                std::memcpy(row, c->bits_default.data(), (size_t)c->bits_row_bytes);
                for (size_t i = 0; i < sc.n_dec; ++i) {
                    const Decoded& d = sc.dec[i];
                    const int64_t o = c->bit_off[d.node];
                    clear_bits(row, o, c->card[d.node]);
                    for (int32_t b : d.bins) row[(o + b) >> 3] |= (uint8_t)(1u << ((o + b) & 7));
                }
            } else if (k == KIND_DENSE && ws) {
                // one run per constrained column: the states from the first to the last non-zero weight (zeros inside are sent);
                // the device expansion reproduces the DENSE row bit for bit (decode.dense_to_wsparse is the Python twin)
                const size_t at = pt.ws.size();
                std::sort(sc.dec.begin(), sc.dec.begin() + (long)sc.n_dec, [](const Decoded& x, const Decoded& y) { return x.node < y.node; });
                for (size_t i = 0; i < sc.n_dec; ++i) {   // runs in column order, like the Python packer
                    const Decoded& d = sc.dec[i];
                    const int card = c->card[d.node];
                    sc.seg.assign((size_t)card, 0.0);
                    for (size_t j = 0; j < d.bins.size(); ++j) sc.seg[d.bins[j]] += d.wts[j];
                    int first = 0, last = -1;
                    for (int s2 = 0; s2 < card; ++s2)
                        if ((float)sc.seg[s2] != 0.f) { if (last < 0) first = s2; last = s2; }
                    bool all_one = true;
                    for (int s2 = 0; s2 < card && all_one; ++s2) all_one = (float)sc.seg[s2] == 1.f;
                    if (all_one) continue;   // the column's segment equals the default: unconstrained, no run
                    int count = last - first + 1;
                    if (last < 0) { first = 0; count = 0; }
                    int done = 0;
                    do {   // the run length is an 8-bit field: 255 states, then continuation runs
                        const int n = std::min(count - done, 255);
                        pt.ws.push_back((uint32_t)d.node | (done ? (1u << 15) : 0u) | ((uint32_t)(first + done) << 16) | ((uint32_t)n << 24));
                        for (int j = 0; j < n; ++j) {
                            const float w = (float)sc.seg[first + done + j];
                            uint32_t bits32;
                            std::memcpy(&bits32, &w, 4);
                            pt.ws.push_back(bits32);
                        }
                        done += n;
                    } while (done < count);
                }
                pt.ws_len.push_back((uint32_t)(pt.ws.size() - at));
                pt.index.push_back((uint32_t)q);
            } else if (k == KIND_DENSE) {
                const size_t at = pt.dense.size();
                pt.dense.resize(at + width);
                float* row = pt.dense.data() + at;
                std::memcpy(row, c->dense_default.data(), width * sizeof(float));
                for (size_t i = 0; i < sc.n_dec; ++i) {
                    const Decoded& d = sc.dec[i];
                    const int card = c->card[d.node];
                    sc.seg.assign((size_t)card, 0.0);
                    for (size_t j = 0; j < d.bins.size(); ++j) sc.seg[d.bins[j]] += d.wts[j];  // np.add.at
                    for (int s2 = 0; s2 < card; ++s2) row[c->dense_off[d.node] + s2] = (float)sc.seg[s2];
                }
                pt.index.push_back((uint32_t)q);
            }
            kind[q] = (uint8_t)k;
        }
    };
    if (n_thr == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_thr; ++t) pool.emplace_back(work, t);
        for (std::thread& th : pool) th.join();
    }
    if (ws) {
        size_t nd = 0, nw = 0;
        for (const Part& pt : part) {
            size_t at = 0;
            for (size_t i = 0; i < pt.index.size(); ++i) {
                const size_t len = pt.ws_len[i];
                if (nd >= dense_capacity || nw + len > ws->capacity || !dense_index) {
                    kind[pt.index[i]] = (uint8_t)KIND_OVERFLOW;
                } else {
                    ws->row_off[nd] = (uint32_t)nw;
                    std::memcpy(ws->words + nw, pt.ws.data() + at, len * 4);
                    nw += len;
                    dense_index[nd++] = pt.index[i];
                }
                at += len;
            }
        }
        ws->row_off[nd] = (uint32_t)nw;
        *ws->n_words = nw;
        *n_dense = nd;
        return BC_OK;
    }
    size_t nd = 0;
    for (const Part& pt : part)
        for (size_t i = 0; i < pt.index.size(); ++i) {
            if (nd >= dense_capacity || !dense_rows || !dense_index) {
                kind[pt.index[i]] = (uint8_t)KIND_OVERFLOW;
                continue;
            }
            std::memcpy(dense_rows + nd * width, pt.dense.data() + i * width, width * sizeof(float));
            dense_index[nd++] = pt.index[i];
        }
    *n_dense = nd;
    return BC_OK;
}
}  // namespace

extern "C" {

int bc_sqlc_compile(const bc_sqlc* c, size_t n_queries, const char* const* sql, uint8_t* kind, void* bits_rows,
                    float* dense_rows, size_t dense_capacity, uint32_t* dense_index, size_t* n_dense) {
    if (!c || (n_queries && (!sql || !kind || !bits_rows)) || !n_dense) {
        bc_set_error("bc_sqlc_compile: bad arguments");
        return BC_EINVAL;
    }
    return run_batch(c, n_queries, [&](size_t q, Scratch& sc) { return sql[q] ? compile_one(*c, sv(sql[q]), sc) : (int)KIND_PYTHON; }, kind,
                     bits_rows, dense_rows, dense_capacity, dense_index, n_dense);
}

int bc_sqlc_set_null(bc_sqlc* c, const char* name, double null_value) {
    if (!c || !name) { bc_set_error("bc_sqlc_set_null: bad arguments"); return BC_EINVAL; }
    auto it = c->by_name.find(std::string_view(name));
    if (it == c->by_name.end()) { bc_set_error("bc_sqlc_set_null: unknown column %s", name); return BC_EINVAL; }
    c->cols[it->second].has_null = true;
    c->cols[it->second].null_num = null_value;
    return BC_OK;
}

int bc_sqlc_column_index(const bc_sqlc* c, const char* name) {
    if (!c || !name) return -1;
    auto it = c->by_name.find(std::string_view(name));
    return it == c->by_name.end() ? -1 : it->second;
}

int bc_sqlc_compile_factors(const bc_sqlc* c, size_t n_factors, const uint32_t* ids, const uint32_t* pred_off, const int32_t* pred_col,
                            const uint8_t* pred_kind, const double* pred_a, const double* pred_b, const uint32_t* fan_mask, uint8_t* kind,
                            void* bits_rows, float* dense_rows, size_t dense_capacity, uint32_t* dense_index, size_t* n_dense,
                            uint32_t* ws_row_off, uint32_t* ws_words, size_t ws_capacity, size_t* n_ws_words) {
    if (!c || (n_factors && (!pred_off || !kind || !bits_rows)) || !n_dense || (ws_words && (!ws_row_off || !n_ws_words))) {
        bc_set_error("bc_sqlc_compile_factors: bad arguments");
        return BC_EINVAL;
    }
    WsOut ws{ws_row_off, ws_words, ws_capacity, n_ws_words};
    return run_batch(c, n_factors,
                     [&](size_t q, Scratch& sc) {
                         const size_t f = ids ? ids[q] : q;
                         return compile_factor(*c, pred_col, pred_kind, pred_a, pred_b, pred_off[f], pred_off[f + 1],
                                               fan_mask != nullptr && fan_mask[f] != 0, sc);
                     },
                     kind, bits_rows, dense_rows, dense_capacity, dense_index, n_dense, ws_words ? &ws : nullptr);
}

}  // extern "C"
