// K-spec code generator: turns ONE learned tree into straight-line sm_100a CUDA.
//
// The reference's docstring promises exactly this -- "Compiles a ppl program into a fixed linear
// algebra program to speed up the inference" (Pgmpy/inference/ExactInference.py:113-114) -- but
// re-plans every query in Python (copy.deepcopy per node, :100).  Here the plan is fixed at model
// load: one THREAD evaluates one query; every message vector lives in registers; every non-zero
// CPT entry T_v[c][p] becomes the immediate operand of one FFMA; there is no shared memory, no
// synchronisation and no data-dependent branch, so all warps run the same instruction stream.
//
//   m_v[p]   = sum_c u_v[c] * T_v[c][p]              (one FFMA per non-zero entry)
//   u_v[c]   = w_v[c] * prod_{children k} m_k[c]     (registers)
//   result   = sum_c u_0[c] * T_0[c]
//
// Children are evaluated largest-subtree first so that the fewest message vectors are live at
// once (Sethi-Ullman order).  Exact zeros of the CPT (30 % of the shipped DMV / Census entries)
// emit no instruction; the number of FFMAs actually emitted is reported in the header comment of
// the generated source and is what the roofline accounting uses.
#include <algorithm>
#include <cstring>
#include <functional>
#include <string>

#include "bc_internal.h"

namespace {

std::string flit(float x) {
    char b[64];
    snprintf(b, sizeof(b), "%.9g", (double)x);
    std::string s(b);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s + "f";
}

struct Gen {
    const bc_model& m;
    std::vector<std::vector<int>> kids;
    std::vector<long long> subtree;
    std::string out;
    long long n_ffma = 0, n_elem = 0;
    bool dense;
    bool any_fan = false;

    explicit Gen(const bc_model& mm, bool d) : m(mm), kids(mm.n), subtree(mm.n, 0), dense(d) {
        for (int v = 1; v < m.n; ++v) kids[m.nodes[v].parent].push_back(v);
        for (int v = m.n - 1; v >= 0; --v) {
            subtree[v] += m.nodes[v].card;
            if (v > 0) subtree[m.nodes[v].parent] += subtree[v];
            if (m.nodes[v].fan_off >= 0) any_fan = true;
        }
        for (int v = 0; v < m.n; ++v)
            std::stable_sort(kids[v].begin(), kids[v].end(), [&](int a, int b) { return subtree[a] > subtree[b]; });
    }

    void line(const std::string& s) { out += s; out += '\n'; }
    static std::string I(long long x) { return std::to_string(x); }

    // selection predicate of state c of node v (range formats) -> C expression of type bool
    void emit_selector(int v) {
        const BcNodeRec& nd = m.nodes[v];
        const int byte = 2 * v, w = byte / 4, sh = (byte % 4) * 8;
        line("    const unsigned lo" + I(v) + " = (d" + I(w) + " >> " + I(sh) + ") & 0xffu, hi" + I(v) + " = min((d" +
             I(w) + " >> " + I(sh + 8) + ") & 0xffu, " + I(nd.card - 1) + "u);");
        if (nd.card <= 32) {
            line("    const unsigned mk" + I(v) + " = hi" + I(v) + " >= lo" + I(v) + " ? ((0xffffffffu >> (31u - hi" +
                 I(v) + ")) & (0xffffffffu << lo" + I(v) + ")) : 0u;");
        } else if (nd.card <= 64) {
            line("    const unsigned long long mk" + I(v) + " = hi" + I(v) + " >= lo" + I(v) +
                 " ? ((0xffffffffffffffffull >> (63u - hi" + I(v) + ")) & (0xffffffffffffffffull << lo" + I(v) +
                 ")) : 0ull;");
        } else {
            line("    const unsigned ln" + I(v) + " = hi" + I(v) + " >= lo" + I(v) + " ? hi" + I(v) + " - lo" + I(v) +
                 " + 1u : 0u;");
        }
    }
    std::string sel(int v, int c) const {
        const BcNodeRec& nd = m.nodes[v];
        if (nd.card <= 32) return "(mk" + I(v) + " & " + I(1u << c) + "u)";
        if (nd.card <= 64) return "(mk" + I(v) + " & " + I(1ull << c) + "ull)";
        return "((" + I(c) + "u - lo" + I(v) + ") < ln" + I(v) + ")";
    }

    // Emits code that leaves the message of node v in registers m<v>_<p>, p < card(parent).
    void emit_message(int v) {
        const BcNodeRec& nd = m.nodes[v];
        const bool root = v == 0;
        line("    // ---- node " + I(v) + ": card " + I(nd.card) + (root ? " (root)" : ", parent " + I(nd.parent)) +
             ", " + I((long long)kids[v].size()) + " children");
        // children first; their messages are indexed by this node's states
        bool have_lam = false;
        for (int k : kids[v]) {
            emit_message(k);
            if (!have_lam) {
                for (int c = 0; c < nd.card; ++c) line("    float l" + I(v) + "_" + I(c) + " = m" + I(k) + "_" + I(c) + ";");
                have_lam = true;
            } else {
                for (int c = 0; c < nd.card; ++c) line("    l" + I(v) + "_" + I(c) + " *= m" + I(k) + "_" + I(c) + ";");
            }
        }
        const bool fan = nd.fan_off >= 0;
        if (!dense) emit_selector(v);
        if (fan) line("    const bool fb" + I(v) + " = (fm" + I(v / 32) + " >> " + I(v % 32) + ") & 1u;");
        const int cols = root ? 1 : nd.card_pa;
        if (root) line("    float r = 0.f;");
        else
            for (int p = 0; p < cols; ++p) line("    float m" + I(v) + "_" + I(p) + " = 0.f;");
        if (dense) {
            // weights of this node: round_up(card,4)/4 float4 loads from the query's row
            for (int j = 0; j < (nd.card + 3) / 4; ++j)
                line("    const float4 w" + I(v) + "_" + I(j) + " = __ldg(row + " + I(nd.lam_off / 4 + j) + ");");
        }
        for (int c = 0; c < nd.card; ++c) {
            ++n_elem;
            std::string base = have_lam ? "l" + I(v) + "_" + I(c) : std::string();
            std::string u;
            if (dense) {
                static const char* comp[4] = {".x", ".y", ".z", ".w"};
                u = "w" + I(v) + "_" + I(c / 4) + comp[c % 4];
                if (fan) u = "(fb" + I(v) + " ? " + u + " * " + flit(m.fan[nd.fan_off + c]) + " : " + u + ")";
                if (have_lam) u = u + " * " + base;
            } else {
                std::string val = have_lam ? base : "1.0f";
                if (fan) {
                    std::string fw = "(fb" + I(v) + " ? " + flit(m.fan[nd.fan_off + c]) + " : 1.0f)";
                    val = have_lam ? base + " * " + fw : fw;
                }
                u = sel(v, c) + " ? " + val + " : 0.f";
            }
            line("    { const float u = " + u + ";");
            const float* T = m.arena.data() + nd.cpt_off;
            if (root) {
                if (T[c] != 0.f) { line("      r = fmaf(u, " + flit(T[c]) + ", r);"); ++n_ffma; }
            } else {
                const float* rowp = T + (size_t)c * nd.stride;
                for (int p = 0; p < cols; ++p) {
                    if (rowp[p] == 0.f) continue;
                    line("      m" + I(v) + "_" + I(p) + " = fmaf(u, " + flit(rowp[p]) + ", m" + I(v) + "_" + I(p) + ");");
                    ++n_ffma;
                }
            }
            line("    }");
        }
    }
};

std::string gen_kernel(const bc_model& m, bool dense, int threads, int min_blocks, long long* n_ffma) {
    Gen g(m, dense);
    g.emit_message(0);
    if (n_ffma) *n_ffma = g.n_ffma;
    std::string s;
    const char* name = dense ? "bc_spec_dense" : "bc_spec_range8";
    s += "// FFMA emitted: " + std::to_string(g.n_ffma) + "  weight elements: " + std::to_string(g.n_elem) + "\n";
    s += std::string("extern \"C\" __global__ void __launch_bounds__(") + std::to_string(threads) + ", " +
         std::to_string(min_blocks) + ") " + name +
         "(const unsigned char* __restrict__ desc, unsigned long long stride, const unsigned* __restrict__ fmask, "
         "float* __restrict__ out, unsigned long long nq)\n{\n";
    s += "  const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;\n";
    s += "  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += step) {\n";
    if (dense) {
        s += "    const float4* __restrict__ row = reinterpret_cast<const float4*>(desc + q * stride);\n";
    } else {
        const int words = (int)(bc_round_up(2LL * m.n, 4) / 4);
        s += "    const unsigned* __restrict__ dw = reinterpret_cast<const unsigned*>(desc + q * stride);\n";
        for (int w = 0; w < words; ++w) s += "    const unsigned d" + std::to_string(w) + " = __ldg(dw + " + std::to_string(w) + ");\n";
    }
    if (g.any_fan) {
        for (int w = 0; w < m.mask_words; ++w)
            s += "    const unsigned fm" + std::to_string(w) + " = fmask ? __ldg(fmask + q * " + std::to_string(m.mask_words) +
                 "ull + " + std::to_string(w) + ") : 0u;\n";
    }
    s += g.out;
    s += "    out[q] = r;\n  }\n}\n\n";
    return s;
}

}  // namespace

// Thread count / occupancy target of the generated kernels.  Register need grows with the widest
// pair of (node, parent) domains on a root-to-leaf path; 128 threads x up to 255 registers always
// fits, small trees get more resident warps.
void bc_spec_geometry(const bc_model& m, int* threads, int* min_blocks) {
    int widest = 0;
    for (int v = 1; v < m.n; ++v) widest = std::max(widest, m.nodes[v].card + m.nodes[v].card_pa);
    *threads = 128;
    if (widest <= 40) *min_blocks = 4;        // <= 128 regs/thread
    else if (widest <= 72) *min_blocks = 3;   // <= 168
    else *min_blocks = 2;                     // <= 255
}

std::string bc_spec_generate(const bc_model& m) {
    int threads, min_blocks;
    bc_spec_geometry(m, &threads, &min_blocks);
    long long f1 = 0, f2 = 0;
    std::string k1 = gen_kernel(m, false, threads, min_blocks, &f1);
    std::string k2 = gen_kernel(m, true, threads, min_blocks, &f2);
    std::string s;
    s += "// Generated by bayescard_b200 spec_codegen (version " + std::to_string(BC_CODEGEN_VERSION) + ") -- do not edit.\n";
    s += "// nodes: " + std::to_string(m.n) + "  dense flop/query: " + std::to_string(m.flops_dense) +
         "  executed flop/query: " + std::to_string(2 * f1) + "\n";
    s += "// BC_SPEC_THREADS=" + std::to_string(threads) + " BC_SPEC_MIN_BLOCKS=" + std::to_string(min_blocks) +
         " BC_SPEC_FFMA=" + std::to_string(f1) + "\n";
    if (m.max_card <= 256) s += k1;
    s += k2;
    return s;
}

uint64_t bc_spec_hash_of(const bc_model& m) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    const int ver = BC_CODEGEN_VERSION;
    mix(&ver, sizeof(ver));
    mix(&m.n, sizeof(m.n));
    for (const BcNodeRec& r : m.nodes) mix(&r, sizeof(r));
    mix(m.arena.data(), m.arena.size() * sizeof(float));
    mix(m.fan.data(), m.fan.size() * sizeof(float));
    return h;
}
