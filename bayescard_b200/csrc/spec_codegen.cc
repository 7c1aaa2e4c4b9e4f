// K-spec code generator: turns ONE learned tree into straight-line sm_100a PTX.
//
// The reference's docstring promises exactly this -- "Compiles a ppl program into a fixed linear
// algebra program to speed up the inference" (Pgmpy/inference/ExactInference.py:113-114) -- but
// re-plans every query in Python (copy.deepcopy per node, :100).  Here the plan is fixed at model
// load: one THREAD evaluates one query; every message vector lives in registers; every non-zero
// CPT entry T_v[c][p] is the 32-bit immediate of exactly one FMA-pipe instruction; there is no
// shared memory, no synchronisation and no data-dependent branch.
//
//   m_v[p]   = sum_c w_v[c] * lambda_v[c] * T_v[c][p]
//   lambda_v = prod_{children k} m_k                       (registers)
//   result   = sum_c w_0[c] * lambda_0[c] * T_0[c]
//
// Two entry points are generated (PTX, so that instruction selection is ours, not a C compiler's):
//
//   bc_spec_bits   w is a BITS row.  The state's bit becomes a PREDICATE (ptxas packs seven of them
//                  per R2P) and guards the FMAs of that state directly:
//                      @p fma.rn.f32 m_p, lambda_c, T, m_p        (leaf:  @p add.f32 m_p, m_p, T)
//                  so selecting costs ~1/7 ALU instruction per state instead of one select per
//                  state.  The first state of every node seeds the accumulators through one select
//                  + multiplies, which removes the zero-initialisation.  On B200 an ALU-pipe
//                  instruction costs two issue slots and an FMA-pipe one costs one (measured,
//                  profiles/r1_microbench_pipes.txt), hence the care.
//   bc_spec_dense  w is a DENSE_F32 row (fractional n_distinct weights): u = w * lambda, plain FMAs.
//
// Children are evaluated largest-subtree first so that the fewest message vectors are live at
// once (Sethi-Ullman order).  Exact zeros of the CPT (30 % of the shipped DMV / Census entries)
// emit no instruction; the number of FMAs actually emitted is reported in the header comment of
// the generated source (BC_SPEC_FFMA) and is what the roofline accounting uses.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "bc_internal.h"

namespace {

std::string fhex(float x) {
    uint32_t b;
    std::memcpy(&b, &x, 4);
    char s[16];
    snprintf(s, sizeof(s), "0f%08X", b);
    return s;
}
inline std::string I(long long x) { return std::to_string(x); }

// Experiment knobs (they are part of the image hash, so every variant has its own cache entry):
//   BC_SPEC_SYNC_EVERY=N   re-align the warps of a CTA with bar.sync every N generated instructions.  The code of
//                          a big tree is a straight line of 100-400 KB, larger than the SM's instruction cache;
//                          warps that drift apart each stream it from L2 on their own, warps kept within one
//                          cache-sized window of each other share the fetched lines.
//   BC_SPEC_DYNAMIC=0      back to the static split.  By DEFAULT the CTAs take their rounds of ntid queries from a global
//                          counter (p_ctr) instead of a fixed stride: SMs do not all run this fetch-bound code at the
//                          same speed (ncu: 510-679 k active cycles per SM on DMV), and with a static split the
//                          slowest SM sets the time.  Measured (profiles/r1_spec_dynamic.txt): DMV 3.00 -> 3.55e9 q/s
//                          (63 -> 75 % of the FP32 peak), Census 9.8 -> 10.8e9 (60 -> 66 %), IMDB unchanged.
//   BC_SPEC_THREADS / BC_SPEC_MIN_BLOCKS   override the CTA geometry chosen by bc_spec_geometry.
int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    if (!e || !*e) return dflt;
    const int v = std::atoi(e);
    return v >= 0 ? v : dflt;
}

struct Gen {
    const bc_model& m;
    const bool dense;
    std::vector<std::vector<int>> kids;
    std::string out;
    long long n_fma = 0;
    int nf = 0, np = 0, nr = 0;
    bool any_fan = false;
    std::vector<std::string> words;  // BITS row words / (dense: unused)
    std::vector<std::string> fmw;    // fan-out mask words

    Gen(const bc_model& mm, bool d) : m(mm), dense(d), kids(mm.n) {
        std::vector<long long> subtree(m.n, 0);
        for (int v = 1; v < m.n; ++v) kids[m.nodes[v].parent].push_back(v);
        for (int v = m.n - 1; v >= 0; --v) {
            subtree[v] += m.nodes[v].card;
            if (v > 0) subtree[m.nodes[v].parent] += subtree[v];
            if (m.nodes[v].fan_off >= 0) any_fan = true;
        }
        for (int v = 0; v < m.n; ++v)
            std::stable_sort(kids[v].begin(), kids[v].end(), [&](int a, int b) { return subtree[a] > subtree[b]; });
    }

    std::string F() { return "%f" + I(++nf); }
    std::string P() { return "%p" + I(++np); }
    std::string R() { return "%r" + I(++nr); }
    int sync_every = 0, since_sync = 0;
    void line(const std::string& s) {
        out += "    "; out += s; out += '\n';
        if (sync_every > 0 && ++since_sync >= sync_every && s[0] != '/') {
            out += "    bar.sync 0;\n";
            since_sync = 0;
        }
    }

    std::string bit_pred(int v, int c) {
        const int b = m.bits[v].bit_off + c;
        std::string t = R(), p = P();
        line("and.b32 " + t + ", " + words[b >> 5] + ", " + I(1LL << (b & 31)) + ";");
        line("setp.ne.u32 " + p + ", " + t + ", 0;");
        return p;
    }

    // Emits the message of node v; returns its registers, one per parent state (root: one).
    std::vector<std::string> message(int v) {
        const BcNodeRec& nd = m.nodes[v];
        const bool root = v == 0;
        const int card = nd.card, cols = root ? 1 : nd.card_pa;
        line("// ---- node " + I(v) + ": card " + I(card) + (root ? " (root)" : ", parent " + I(nd.parent)));
        std::vector<std::string> lam;
        for (int k : kids[v]) {
            std::vector<std::string> mk = message(k);
            if (lam.empty()) lam = mk;
            else
                for (int c = 0; c < card; ++c) line("mul.f32 " + lam[c] + ", " + lam[c] + ", " + mk[c] + ";");
        }
        const bool fan = nd.fan_off >= 0;
        std::string pfb;
        if (fan) {
            std::string t = R();
            pfb = P();
            line("and.b32 " + t + ", " + fmw[v >> 5] + ", " + I(1LL << (v & 31)) + ";");
            line("setp.ne.u32 " + pfb + ", " + t + ", 0;");
        }
        const float* T = m.arena.data() + nd.cpt_off;
        const int stride = root ? 1 : nd.stride;
        auto Tat = [&](int c, int p) { return root ? T[c] : T[(size_t)c * stride + p]; };
        // the seed row: the state with the most non-zero entries (fewest accumulators left to zero)
        int seed = 0, best = -1;
        for (int c = 0; c < card; ++c) {
            int nz = 0;
            for (int p = 0; p < cols; ++p) nz += Tat(c, p) != 0.f;
            if (nz > best) { best = nz; seed = c; }
        }
        std::vector<std::string> acc(cols);
        std::vector<int> order;
        order.push_back(seed);
        for (int c = 0; c < card; ++c)
            if (c != seed) order.push_back(c);
        // dense rows: the weights of this node are loaded four at a time, right before the first state of the group
        // is folded in (loading the whole node up front keeps up to card more registers live and spills on IMDB)
        std::vector<std::string> wreg(dense ? ((card + 3) / 4) * 4 : 0);
        auto need_weight = [&](int c) {
            const int j = c / 4;
            if (!wreg[4 * j].empty()) return;
            std::string a = F(), b = F(), c2 = F(), d = F();
            line("ld.global.nc.v4.f32 {" + a + ", " + b + ", " + c2 + ", " + d + "}, [%rrow+" + I(4LL * (nd.lam_off + 4 * j)) + "];");
            wreg[4 * j] = a; wreg[4 * j + 1] = b; wreg[4 * j + 2] = c2; wreg[4 * j + 3] = d;
        };
        for (int c : order) {
            bool any = false;
            for (int p = 0; p < cols; ++p) any |= Tat(c, p) != 0.f;
            if (!any && c != seed) continue;
            // lambda of this state, times the fan-out value when the query's fan bit is set
            std::string lc = lam.empty() ? std::string() : lam[c];
            if (fan) {
                const std::string fv = fhex(m.fan[nd.fan_off + c]);
                if (lc.empty()) {
                    lc = F();
                    line("selp.f32 " + lc + ", " + fv + ", 0f3F800000, " + pfb + ";");
                } else {
                    line("@" + pfb + " mul.f32 " + lc + ", " + lc + ", " + fv + ";");
                }
            }
            if (dense) {
                need_weight(c);
                std::string u = wreg[c];
                if (!lc.empty()) {
                    u = F();
                    line("mul.f32 " + u + ", " + wreg[c] + ", " + lc + ";");
                }
                for (int p = 0; p < cols; ++p) {
                    const float t = Tat(c, p);
                    if (c == seed) {
                        acc[p] = F();
                        if (t != 0.f) { line("mul.f32 " + acc[p] + ", " + u + ", " + fhex(t) + ";"); ++n_fma; }
                        else line("mov.f32 " + acc[p] + ", 0f00000000;");
                    } else if (t != 0.f) {
                        line("fma.rn.f32 " + acc[p] + ", " + u + ", " + fhex(t) + ", " + acc[p] + ";");
                        ++n_fma;
                    }
                }
                continue;
            }
            const std::string pc = bit_pred(v, c);
            if (c == seed) {
                std::string u = F();
                line("selp.f32 " + u + ", " + (lc.empty() ? std::string("0f3F800000") : lc) + ", 0f00000000, " + pc + ";");
                for (int p = 0; p < cols; ++p) {
                    const float t = Tat(c, p);
                    acc[p] = F();
                    if (t != 0.f) { line("mul.f32 " + acc[p] + ", " + u + ", " + fhex(t) + ";"); ++n_fma; }
                    else line("mov.f32 " + acc[p] + ", 0f00000000;");
                }
            } else {
                for (int p = 0; p < cols; ++p) {
                    const float t = Tat(c, p);
                    if (t == 0.f) continue;
                    ++n_fma;
                    if (lc.empty()) line("@" + pc + " add.f32 " + acc[p] + ", " + acc[p] + ", " + fhex(t) + ";");
                    else line("@" + pc + " fma.rn.f32 " + acc[p] + ", " + lc + ", " + fhex(t) + ", " + acc[p] + ";");
                }
            }
        }
        return acc;
    }
};

std::string gen_kernel(const bc_model& m, bool dense, int threads, int min_blocks, long long* n_fma) {
    Gen g(m, dense);
    const int sync_every = env_int("BC_SPEC_SYNC_EVERY", 0);
    g.sync_every = sync_every;
    const int words = m.bits_words;
    if (!dense)
        for (int w = 0; w < words; ++w) g.words.push_back(g.R());
    for (int w = 0; w < m.mask_words; ++w) g.fmw.push_back(g.R());
    const std::string res = g.message(0)[0];
    if (n_fma) *n_fma = g.n_fma;

    const bool dynamic = env_int("BC_SPEC_DYNAMIC", 1) != 0;
    std::string s;
    const char* name = dense ? "bc_spec_dense" : "bc_spec_bits";
    s += std::string(".visible .entry ") + name +
         "(\n    .param .u64 p_desc,\n    .param .u64 p_stride,\n    .param .u64 p_fmask,\n    .param .u64 p_out,\n"
         "    .param .u64 p_nq,\n    .param .u64 p_ctr\n)\n.maxntid " + I(threads) + ", 1, 1\n.minnctapersm " + I(min_blocks) + "\n{\n";
    s += "    .reg .pred %p<" + I(g.np + 1) + ">;\n    .reg .pred %pfm, %pdone;\n";
    s += "    .reg .f32 %f<" + I(g.nf + 1) + ">;\n    .reg .b32 %r<" + I(g.nr + 1) + ">;\n";
    s += "    .reg .b32 %t0, %t1, %t2, %t3, %t4;\n";
    s += "    .reg .b64 %rdesc, %rstride, %rfmask, %rout, %rnq, %rq, %rstep, %rtmp, %rrow, %rqb, %rqc;\n";
    s += "    .reg .b64 %rctr, %rnext, %rnt, %rtid;\n";
    s += "    .reg .pred %pvalid, %pt0, %plast;\n";
    if (dynamic) s += "    .shared .align 8 .b64 bc_slot[2];\n";
    auto e = [&](const std::string& x) { s += "    " + x + "\n"; };
    e("ld.param.u64 %rdesc, [p_desc];");
    e("ld.param.u64 %rstride, [p_stride];");
    e("ld.param.u64 %rfmask, [p_fmask];");
    e("ld.param.u64 %rout, [p_out];");
    e("ld.param.u64 %rnq, [p_nq];");
    e("cvta.to.global.u64 %rdesc, %rdesc;");
    e("cvta.to.global.u64 %rout, %rout;");
    e("setp.ne.u64 %pfm, %rfmask, 0;");
    e("@%pfm cvta.to.global.u64 %rfmask, %rfmask;");
    const bool clamp = dynamic || sync_every > 0;  // threads past the end recompute the last query and skip the store
    if (dynamic) {
        // Round k of a CTA = ntid consecutive queries starting at the value its thread 0 drew from the global
        // counter; the draw for round k+1 is issued at the start of round k-1 (its latency hides behind a whole
        // round) and handed over through a two-slot shared-memory mailbox, one bar.sync per round.
        e("ld.param.u64 %rctr, [p_ctr];");
        e("cvta.to.global.u64 %rctr, %rctr;");
        e("mov.u32 %t0, %tid.x;");
        e("mov.u32 %t2, %ntid.x;");
        e("cvt.u64.u32 %rtid, %t0;");
        e("cvt.u64.u32 %rnt, %t2;");
        e("setp.eq.u32 %pt0, %t0, 0;");
        e("mov.u32 %t4, bc_slot;");
        e("@%pt0 atom.global.add.u64 %rnext, [%rctr], %rnt;");
        e("@%pt0 st.shared.u64 [%t4], %rnext;");
        e("@%pt0 atom.global.add.u64 %rnext, [%rctr], %rnt;");
        s += "LOOP:\n";
        e("bar.sync 0;");
        e("ld.shared.u64 %rqb, [%t4];");
        e("setp.ge.u64 %pdone, %rqb, %rnq;");
        e("@%pdone bra DONE;");
        e("mov.u32 %t1, bc_slot;");
        e("sub.u32 %t3, %t4, %t1;");
        e("xor.b32 %t3, %t3, 8;");
        e("add.u32 %t4, %t1, %t3;");
        e("@%pt0 st.shared.u64 [%t4], %rnext;");
        e("@%pt0 atom.global.add.u64 %rnext, [%rctr], %rnt;");
        e("add.u64 %rq, %rqb, %rtid;");
    } else {
        // query index of round k:  ((k * warps_per_cta + warp) * n_cta + cta) * 32 + lane.
        // Consecutive warp slots belong to DIFFERENT CTAs, so the last, partial round is spread over all
        // CTAs (and SMs) instead of filling the first ones only.
        e("mov.u32 %t0, %tid.x;");
        e("mov.u32 %t1, %ctaid.x;");
        e("mov.u32 %t2, %ntid.x;");
        e("mov.u32 %t3, %nctaid.x;");
        e("shr.u32 %t4, %t0, 5;");
        e("mad.lo.u32 %t4, %t4, %t3, %t1;");
        e("and.b32 %t0, %t0, 31;");
        e("mul.wide.u32 %rq, %t4, 32;");
        e("cvt.u64.u32 %rtmp, %t0;");
        e("add.u64 %rq, %rq, %rtmp;");
        e("mul.wide.u32 %rstep, %t2, %t3;");
        if (sync_every > 0) e("mul.wide.u32 %rqb, %t1, 32;");  // first query of the CTA's first warp slot
        s += "LOOP:\n";
        if (sync_every > 0) {
            // bar.sync needs a CTA-uniform trip count: loop while the CTA's FIRST slot is in range
            e("setp.ge.u64 %pdone, %rqb, %rnq;");
            e("@%pdone bra DONE;");
        } else {
            e("setp.ge.u64 %pdone, %rq, %rnq;");
            e("@%pdone bra DONE;");
        }
    }
    if (clamp) {
        e("setp.lt.u64 %pvalid, %rq, %rnq;");
        e("sub.u64 %rqc, %rnq, 1;");
        e("min.u64 %rqc, %rqc, %rq;");
        e("mad.lo.u64 %rrow, %rqc, %rstride, %rdesc;");
    } else {
        e("mad.lo.u64 %rrow, %rq, %rstride, %rdesc;");
    }
    if (!dense)
        for (int w = 0; w < words; w += 4)
            e("ld.global.nc.v4.u32 {" + g.words[w] + ", " + g.words[w + 1] + ", " + g.words[w + 2] + ", " + g.words[w + 3] +
              "}, [%rrow+" + I(4LL * w) + "];");
    if (g.any_fan) {
        e(std::string("mad.lo.u64 %rtmp, ") + (clamp ? "%rqc" : "%rq") + ", " + I(4LL * m.mask_words) + ", %rfmask;");
        for (int w = 0; w < m.mask_words; ++w) {
            e("mov.u32 " + g.fmw[w] + ", 0;");
            e("@%pfm ld.global.nc.u32 " + g.fmw[w] + ", [%rtmp+" + I(4LL * w) + "];");
        }
    }
    s += g.out;
    e("shl.b64 %rtmp, %rq, 2;");
    e("add.u64 %rtmp, %rtmp, %rout;");
    e(std::string(clamp ? "@%pvalid " : "") + "st.global.f32 [%rtmp], " + res + ";");
    if (!dynamic) {
        e("add.u64 %rq, %rq, %rstep;");
        if (sync_every > 0) e("add.u64 %rqb, %rqb, %rstep;");
    }
    e("bra LOOP;");
    if (dynamic) {
        // The last CTA to leave zeroes the counter for the next launch.  Thread 0 first consumes its outstanding
        // draw (the shift is 0 for any real counter value, but the address now depends on it), so no draw of any
        // CTA can land after the reset.
        s += "DONE:\n";
        e("@!%pt0 bra OUT;");
        e("shr.u64 %rtmp, %rnext, 63;");
        e("add.u64 %rctr, %rctr, %rtmp;");
        e("mov.u32 %t3, %nctaid.x;");
        e("sub.u32 %t3, %t3, 1;");
        e("atom.global.inc.u32 %t1, [%rctr+8], %t3;");
        e("setp.eq.u32 %plast, %t1, %t3;");
        e("mov.b64 %rtmp, 0;");
        e("@%plast st.global.u64 [%rctr], %rtmp;");
        s += "OUT:\n    ret;\n}\n\n";
        return s;
    }
    s += "DONE:\n    ret;\n}\n\n";
    return s;
}

}  // namespace

// Thread count / occupancy target of the generated kernels.  Register need grows with the widest
// (node, parent) pair of domains on a root-to-leaf path.  Measured on B200 (profiles/r1_spec_v2_sweep.txt):
// small trees like 256-thread CTAs with <= 85 registers; DMV-sized ones want all 255 registers and no
// spills; beyond that three resident CTAs with a few spilled values beat two without.
void bc_spec_geometry(const bc_model& m, int* threads, int* min_blocks) {
    int widest = 0;
    for (int v = 1; v < m.n; ++v) widest = std::max(widest, m.nodes[v].card + m.nodes[v].card_pa);
    if (widest <= 40) { *threads = 256; *min_blocks = 3; }
    else if (widest <= 110) { *threads = 128; *min_blocks = 2; }
    else { *threads = 128; *min_blocks = 3; }
    const int t = env_int("BC_SPEC_THREADS", 0), b = env_int("BC_SPEC_MIN_BLOCKS", 0);
    if (t >= 32 && t <= 1024 && t % 32 == 0) *threads = t;
    if (b >= 1 && b <= 16) *min_blocks = b;
}

std::string bc_spec_generate(const bc_model& m) {
    int threads, min_blocks;
    bc_spec_geometry(m, &threads, &min_blocks);
    long long f1 = 0, f2 = 0;
    std::string k1 = gen_kernel(m, false, threads, min_blocks, &f1);
    std::string k2 = gen_kernel(m, true, threads, min_blocks, &f2);
    std::string s;
    s += "//\n// Generated by bayescard_b200 spec_codegen (version " + I(BC_CODEGEN_VERSION) + ") -- do not edit.\n";
    s += "// nodes: " + I(m.n) + "  dense flop/query: " + I(m.flops_dense) + "  executed flop/query: " + I(2 * f1) + "\n";
    s += "// BC_SPEC_THREADS=" + I(threads) + " BC_SPEC_MIN_BLOCKS=" + I(min_blocks) + " BC_SPEC_FFMA=" + I(f1) + "\n//\n";
    s += ".version 8.7\n.target sm_100a\n.address_size 64\n\n";
    // geometry travels with the image: {threads per CTA, queries per thread and trip, codegen version, 0}
    s += ".visible .global .align 4 .u32 bc_spec_meta[4] = {" + I(threads) + ", 1, " + I(BC_CODEGEN_VERSION) + ", 0};\n\n";
    s += k1;
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
    const int knobs[4] = {env_int("BC_SPEC_SYNC_EVERY", 0), env_int("BC_SPEC_THREADS", 0), env_int("BC_SPEC_MIN_BLOCKS", 0),
                          env_int("BC_SPEC_DYNAMIC", 1)};
    if (knobs[0] || knobs[1] || knobs[2] || knobs[3] != 1) mix(knobs, sizeof(knobs));
    mix(&m.n, sizeof(m.n));
    for (const BcNodeRec& r : m.nodes) mix(&r, sizeof(r));
    mix(m.arena.data(), m.arena.size() * sizeof(float));
    mix(m.fan.data(), m.fan.size() * sizeof(float));
    return h;
}
