// Host-side set-up, native: Stim-syntax circuit text -> detector error model (independent error
// mechanisms + graphlike decomposition).  Replaces UPSTREAM
// stim.Circuit.detector_error_model(decompose_errors=True, approximate_disjoint_errors=True), which sinter
// invokes for every task because the reference builds its Tasks without a DEM
// (reference threshold_lab.py:151).
//
// Algorithm (the same as graphqec_b200/dem.py, which stays as the readable statement of it): walk the
// circuit backwards keeping, per qubit, the set of detectors / observables an X or a Z inserted at that
// point would flip; a noise channel emits its independent Pauli components with the symptom sets current
// at its position; equal symptom sets merge with p <- p1 (1 - p2) + p2 (1 - p1), in emission order, so
// the probabilities equal the Python builder's bit for bit.  Symptom sets are sorted id vectors
// (detector d -> d, observable k -> 0x80000000 | k): they stay small (a handful of ids), so a symmetric
// difference is a short merge, where the Python version XORs D-bit integers.
//
// No device code in this file; it is a .cu only so that the library has one build rule.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "gq_common.h"

namespace {

typedef std::vector<uint32_t> Sym;
const uint32_t OBS_BIT = 0x80000000u;

enum Gate : uint8_t {
    G_R, G_RX, G_M, G_MX, G_MR, G_MRX, G_H, G_PAULI, G_S, G_CX, G_CZ, G_SWAP,
    G_DEP1, G_XERR, G_YERR, G_ZERR, G_DEP2, G_DETECTOR, G_OBSINC, G_NOP
};

struct Op {
    Gate gate;
    double arg;
    bool has_arg;
    uint32_t t0, t1;   // targets [t0, t1) in the flat target array (rec targets are negative)
};

struct GateName { const char *name; Gate gate; };
const GateName GATE_NAMES[] = {
    {"R", G_R}, {"RZ", G_R}, {"RX", G_RX}, {"M", G_M}, {"MZ", G_M}, {"MX", G_MX}, {"MR", G_MR}, {"MRZ", G_MR},
    {"MRX", G_MRX}, {"H", G_H}, {"H_XZ", G_H}, {"X", G_PAULI}, {"Y", G_PAULI}, {"Z", G_PAULI}, {"I", G_PAULI},
    {"S", G_S}, {"S_DAG", G_S}, {"CX", G_CX}, {"CNOT", G_CX}, {"ZCX", G_CX}, {"CZ", G_CZ}, {"ZCZ", G_CZ},
    {"SWAP", G_SWAP}, {"DEPOLARIZE1", G_DEP1}, {"X_ERROR", G_XERR}, {"Y_ERROR", G_YERR}, {"Z_ERROR", G_ZERR},
    {"DEPOLARIZE2", G_DEP2}, {"DETECTOR", G_DETECTOR}, {"OBSERVABLE_INCLUDE", G_OBSINC}, {"TICK", G_NOP},
    {"QUBIT_COORDS", G_NOP}, {"SHIFT_COORDS", G_NOP},
};

struct SymHash {
    size_t operator()(const Sym &s) const {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ s.size();
        for (uint32_t x : s) { h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xff51afd7ed558ccdull; }
        return (size_t)h;
    }
};

void sym_xor(Sym &a, const Sym &b, Sym &tmp) {   // a ^= b (sorted symmetric difference)
    if (b.empty()) return;
    tmp.clear();
    size_t i = 0, j = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i] < b[j]) tmp.push_back(a[i++]);
        else if (b[j] < a[i]) tmp.push_back(b[j++]);
        else { ++i; ++j; }
    }
    tmp.insert(tmp.end(), a.begin() + i, a.end());
    tmp.insert(tmp.end(), b.begin() + j, b.end());
    a.swap(tmp);
}

struct Parser {
    std::vector<Op> ops;
    std::vector<int64_t> targets;
    std::string err;

    // one instruction line (no block syntax); returns false on error
    bool line(const char *b, const char *e) {
        while (b < e && isspace((unsigned char)*b)) ++b;
        const char *n0 = b;
        while (b < e && (isalnum((unsigned char)*b) || *b == '_')) ++b;
        if (b == n0) { err = "cannot parse circuit line: " + std::string(n0, e); return false; }
        std::string name(n0, b);
        for (char &c : name) c = (char)toupper((unsigned char)c);
        const GateName *g = nullptr;
        for (const GateName &cand : GATE_NAMES) if (name == cand.name) { g = &cand; break; }
        if (!g) { err = "Gate not supported by the graphqec_b200 circuit IR: '" + name + "'"; return false; }
        Op op;
        op.gate = g->gate; op.arg = 0; op.has_arg = false;
        while (b < e && isspace((unsigned char)*b)) ++b;
        if (b < e && *b == '(') {
            const char *close = (const char *)memchr(b, ')', e - b);
            if (!close) { err = "missing ')' in: " + std::string(n0, e); return false; }
            std::string args(b + 1, close);
            if (args.find_first_not_of(" \t") != std::string::npos) { op.arg = strtod(args.c_str(), nullptr); op.has_arg = true; }
            b = close + 1;
        }
        op.t0 = (uint32_t)targets.size();
        while (b < e) {
            while (b < e && isspace((unsigned char)*b)) ++b;
            if (b >= e) break;
            if (e - b > 4 && !strncmp(b, "rec[", 4)) {
                char *end;
                long v = strtol(b + 4, &end, 10);
                if (end >= e || *end != ']' || v >= 0) { err = "bad rec target in: " + std::string(n0, e); return false; }
                targets.push_back(v);
                b = end + 1;
            } else {
                char *end;
                long v = strtol(b, &end, 10);
                if (end == b || v < 0) { err = "bad target in: " + std::string(n0, e); return false; }
                targets.push_back(v);
                b = end;
            }
        }
        op.t1 = (uint32_t)targets.size();
        if (op.gate != G_NOP) ops.push_back(op);
        return true;
    }

    // program text with REPEAT n { ... } blocks (unrolled)
    bool block(const char *&p, const char *end, bool inner) {
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', end - p);
            const char *le = nl ? nl : end;
            const char *hash = (const char *)memchr(p, '#', le - p);
            const char *b = p, *e = hash ? hash : le;
            p = nl ? nl + 1 : end;
            while (b < e && isspace((unsigned char)*b)) ++b;
            while (e > b && isspace((unsigned char)e[-1])) --e;
            if (b == e) continue;
            if (*b == '}') {
                if (!inner) { err = "unmatched '}'"; return false; }
                return true;
            }
            if (e - b > 6 && !strncmp(b, "REPEAT", 6) && e[-1] == '{') {
                long reps = strtol(b + 6, nullptr, 10);
                if (reps < 0) { err = "bad REPEAT count"; return false; }
                const size_t o0 = ops.size();
                if (!block(p, end, true)) return false;
                const size_t o1 = ops.size();
                if (reps == 0) ops.resize(o0);
                for (long r = 1; r < reps; ++r)
                    for (size_t k = o0; k < o1; ++k) ops.push_back(ops[k]);   // targets are shared: rec offsets are relative
                continue;
            }
            if (!line(b, e)) return false;
        }
        if (inner) { err = "missing '}'"; return false; }
        return true;
    }
};

struct Row { Sym dets; uint64_t obs; double p; };

struct Decomposer {
    // graphlike symptom (1 or 2 detectors) -> its observable masks, ascending
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint64_t>> graphlike;
    std::vector<std::pair<uint32_t, uint32_t>> acc;
    std::vector<uint32_t> best_a, best_b;
    std::vector<uint64_t> best_o;
    bool have = false;
    uint64_t omask = 0;

    void leaf() {
        // first combination (the first block's choice varies slowest) whose masks XOR to omask
        const size_t n = acc.size();
        std::vector<const std::vector<uint64_t> *> ch(n);
        for (size_t i = 0; i < n; ++i) ch[i] = &graphlike[acc[i]];
        std::vector<size_t> idx(n, 0);
        for (;;) {
            uint64_t x = 0;
            for (size_t i = 0; i < n; ++i) x ^= (*ch[i])[idx[i]];
            if (x == omask) {
                have = true;
                best_a.clear(); best_b.clear(); best_o.clear();
                for (size_t i = 0; i < n; ++i) { best_a.push_back(acc[i].first); best_b.push_back(acc[i].second); best_o.push_back((*ch[i])[idx[i]]); }
                return;
            }
            size_t k = n;
            while (k > 0) { --k; if (++idx[k] < ch[k]->size()) break; idx[k] = 0; if (k == 0) return; }
            if (n == 0) return;
        }
    }
    void rec(const Sym &rest) {
        if (have && acc.size() >= best_a.size()) return;
        if (rest.empty()) { leaf(); return; }
        const uint32_t a = rest[0];
        for (size_t j = 1; j < rest.size(); ++j) {
            std::pair<uint32_t, uint32_t> blk(a, rest[j]);
            if (graphlike.count(blk)) {
                Sym nxt;
                nxt.reserve(rest.size() - 2);
                for (size_t k = 1; k < rest.size(); ++k) if (k != j) nxt.push_back(rest[k]);
                acc.push_back(blk);
                rec(nxt);
                acc.pop_back();
            }
        }
        std::pair<uint32_t, uint32_t> single(a, GQ_NO_DET);
        if (graphlike.count(single)) {
            Sym nxt(rest.begin() + 1, rest.end());
            acc.push_back(single);
            rec(nxt);
            acc.pop_back();
        }
    }
};

}  // namespace

struct gq_dem_build {
    uint32_t D = 0, O = 0;
    std::vector<double> p;
    std::vector<uint32_t> det_indptr, det_idx, obs_indptr, obs_idx, comp_indptr, comp_a, comp_b;
    std::vector<uint64_t> comp_obs;
    bool decomposed = false;
};

extern "C" int gq_circuit_to_dem(const char *stim_text, int decompose_errors, gq_dem_build **out) {
    if (!stim_text || !out) GQ_FAIL(GQ_E_INVALID, "gq_circuit_to_dem: NULL argument");
    *out = nullptr;
    Parser P;
    const char *p = stim_text, *end = stim_text + strlen(stim_text);
    if (!P.block(p, end, false)) GQ_FAIL(GQ_E_INVALID, P.err);

    // ---- forward pass: counts, and per measurement the detectors / observables that include it
    uint32_t nq = 0, nmeas = 0, ndet = 0, nobs = 0;
    for (const Op &op : P.ops) {
        const uint32_t nt = op.t1 - op.t0;
        switch (op.gate) {
        case G_M: case G_MX: case G_MR: case G_MRX: nmeas += nt; break;
        case G_DETECTOR: ++ndet; break;
        case G_OBSINC:
            if (!op.has_arg || op.arg < 0 || op.arg > 63) GQ_FAIL(GQ_E_INVALID, "OBSERVABLE_INCLUDE needs an observable index argument in 0..63");
            nobs = std::max(nobs, (uint32_t)op.arg + 1);
            break;
        default: break;
        }
        if (op.gate != G_DETECTOR && op.gate != G_OBSINC) {
            if ((op.gate == G_CX || op.gate == G_CZ || op.gate == G_SWAP || op.gate == G_DEP2) && (nt & 1))
                GQ_FAIL(GQ_E_INVALID, "two-qubit instruction with an odd number of targets");
            for (uint32_t k = op.t0; k < op.t1; ++k) {
                if (P.targets[k] < 0) GQ_FAIL(GQ_E_INVALID, "rec targets are only valid in DETECTOR / OBSERVABLE_INCLUDE");
                nq = std::max<uint32_t>(nq, (uint32_t)P.targets[k] + 1);
            }
        }
    }
    std::vector<Sym> mmask(nmeas);
    {
        uint32_t seen = 0, det = 0;
        for (const Op &op : P.ops) {
            if (op.gate == G_M || op.gate == G_MX || op.gate == G_MR || op.gate == G_MRX) seen += op.t1 - op.t0;
            else if (op.gate == G_DETECTOR || op.gate == G_OBSINC) {
                const uint32_t id = op.gate == G_DETECTOR ? det++ : (OBS_BIT | (uint32_t)op.arg);
                for (uint32_t k = op.t0; k < op.t1; ++k) {
                    const int64_t r = P.targets[k];
                    if (r >= 0) GQ_FAIL(GQ_E_INVALID, "DETECTOR / OBSERVABLE_INCLUDE take rec[-k] targets only");
                    if (-r > (int64_t)seen) GQ_FAIL(GQ_E_INVALID, "rec target refers to a measurement before the start of the circuit");
                    mmask[seen + r].push_back(id);
                }
            }
        }
        for (Sym &s : mmask) {   // an id listed an even number of times cancels
            std::sort(s.begin(), s.end());
            Sym t;
            for (size_t i = 0; i < s.size();) {
                size_t j = i;
                while (j < s.size() && s[j] == s[i]) ++j;
                if ((j - i) & 1) t.push_back(s[i]);
                i = j;
            }
            s.swap(t);
        }
    }

    // ---- backward sensitivity sweep
    std::vector<Sym> xs(nq), zs(nq);
    std::unordered_map<Sym, uint32_t, SymHash> index;   // symptom -> slot in probs (emission order)
    std::vector<double> probs;
    std::vector<const Sym *> keys;
    Sym tmp, acc2;
    auto emit = [&](const Sym &s, double pr) {
        if (s.empty() || pr <= 0.0) return;
        auto it = index.find(s);
        if (it == index.end()) {
            it = index.emplace(s, (uint32_t)probs.size()).first;
            probs.push_back(pr);
            keys.push_back(&it->first);
        } else {
            double &old = probs[it->second];
            old = old * (1.0 - pr) + pr * (1.0 - old);
        }
    };
    uint32_t m = nmeas;
    for (size_t oi = P.ops.size(); oi-- > 0;) {
        const Op &op = P.ops[oi];
        const int64_t *t = P.targets.data();
        switch (op.gate) {
        case G_DEP1: {
            const double pr = op.has_arg ? op.arg : 0.0;
            if (pr > 0.75) GQ_FAIL(GQ_E_INVALID, "DEPOLARIZE1 probability above 3/4 has no independent-component form");
            const double q1 = 0.5 - 0.5 * sqrt(1.0 - 4.0 * pr / 3.0);
            for (uint32_t k = op.t0; k < op.t1; ++k) {
                const uint32_t q = (uint32_t)t[k];
                emit(xs[q], q1);
                emit(zs[q], q1);
                acc2 = xs[q]; sym_xor(acc2, zs[q], tmp);
                emit(acc2, q1);
            }
            break;
        }
        case G_XERR: for (uint32_t k = op.t0; k < op.t1; ++k) emit(xs[t[k]], op.has_arg ? op.arg : 0.0); break;
        case G_ZERR: for (uint32_t k = op.t0; k < op.t1; ++k) emit(zs[t[k]], op.has_arg ? op.arg : 0.0); break;
        case G_YERR:
            for (uint32_t k = op.t0; k < op.t1; ++k) { acc2 = xs[t[k]]; sym_xor(acc2, zs[t[k]], tmp); emit(acc2, op.has_arg ? op.arg : 0.0); }
            break;
        case G_DEP2: {
            const double pr = op.has_arg ? op.arg : 0.0;
            if (pr > 15.0 / 16.0) GQ_FAIL(GQ_E_INVALID, "DEPOLARIZE2 probability above 15/16 has no independent-component form");
            const double q2 = 0.5 - 0.5 * pow(1.0 - 16.0 * pr / 15.0, 0.125);
            for (uint32_t k = op.t0; k < op.t1; k += 2) {
                const Sym *basis[4] = {&xs[t[k]], &zs[t[k]], &xs[t[k + 1]], &zs[t[k + 1]]};
                for (int sel = 1; sel < 16; ++sel) {
                    acc2.clear();
                    for (int b = 0; b < 4; ++b) if (sel & (1 << b)) sym_xor(acc2, *basis[b], tmp);
                    emit(acc2, q2);
                }
            }
            break;
        }
        case G_M: case G_MX:
            for (uint32_t k = op.t1; k-- > op.t0;) {
                --m;
                sym_xor(op.gate == G_M ? xs[t[k]] : zs[t[k]], mmask[m], tmp);
                emit(mmask[m], op.has_arg ? op.arg : 0.0);
            }
            break;
        case G_MR: case G_MRX:
            for (uint32_t k = op.t1; k-- > op.t0;) {
                --m;
                if (op.gate == G_MR) { xs[t[k]] = mmask[m]; zs[t[k]].clear(); }
                else { zs[t[k]] = mmask[m]; xs[t[k]].clear(); }
                emit(mmask[m], op.has_arg ? op.arg : 0.0);
            }
            break;
        case G_R: case G_RX: for (uint32_t k = op.t0; k < op.t1; ++k) { xs[t[k]].clear(); zs[t[k]].clear(); } break;
        case G_H: for (uint32_t k = op.t0; k < op.t1; ++k) xs[t[k]].swap(zs[t[k]]); break;
        case G_S: for (uint32_t k = op.t0; k < op.t1; ++k) sym_xor(xs[t[k]], zs[t[k]], tmp); break;
        case G_CX:
            for (uint32_t k = op.t1; k >= op.t0 + 2; k -= 2) { sym_xor(xs[t[k - 2]], xs[t[k - 1]], tmp); sym_xor(zs[t[k - 1]], zs[t[k - 2]], tmp); }
            break;
        case G_CZ:
            for (uint32_t k = op.t1; k >= op.t0 + 2; k -= 2) { sym_xor(xs[t[k - 2]], zs[t[k - 1]], tmp); sym_xor(xs[t[k - 1]], zs[t[k - 2]], tmp); }
            break;
        case G_SWAP:
            for (uint32_t k = op.t1; k >= op.t0 + 2; k -= 2) { xs[t[k - 2]].swap(xs[t[k - 1]]); zs[t[k - 2]].swap(zs[t[k - 1]]); }
            break;
        default: break;   // Pauli gates and annotations do not move sensitivities
        }
    }
    if (m != 0) GQ_FAIL(GQ_E_INTERNAL, "measurement bookkeeping out of sync");

    // ---- rows sorted by (detector tuple, observable mask)
    std::vector<Row> rows(probs.size());
    for (size_t i = 0; i < probs.size(); ++i) {
        Row &r = rows[i];
        r.p = probs[i]; r.obs = 0;
        for (uint32_t id : *keys[i]) {
            if (id & OBS_BIT) r.obs |= 1ull << (id & 63);
            else r.dets.push_back(id);
        }
    }
    std::sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.dets != b.dets ? a.dets < b.dets : a.obs < b.obs; });

    gq_dem_build *B = new gq_dem_build();
    B->D = ndet; B->O = nobs;
    B->det_indptr.push_back(0); B->obs_indptr.push_back(0); B->comp_indptr.push_back(0);
    for (const Row &r : rows) {
        B->p.push_back(r.p);
        B->det_idx.insert(B->det_idx.end(), r.dets.begin(), r.dets.end());
        B->det_indptr.push_back((uint32_t)B->det_idx.size());
        for (int k = 0; k < 64; ++k) if ((r.obs >> k) & 1) B->obs_idx.push_back((uint32_t)k);
        B->obs_indptr.push_back((uint32_t)B->obs_idx.size());
    }
    B->decomposed = decompose_errors != 0;
    if (decompose_errors) {
        Decomposer dc;
        for (const Row &r : rows)
            if (r.dets.size() == 1 || r.dets.size() == 2)
                dc.graphlike[std::make_pair(r.dets[0], r.dets.size() == 2 ? r.dets[1] : GQ_NO_DET)].push_back(r.obs);   // rows are sorted: masks ascend
        for (const Row &r : rows) {
            if (r.dets.size() <= 2) {
                B->comp_a.push_back(r.dets.size() >= 1 ? r.dets[0] : GQ_NO_DET);
                B->comp_b.push_back(r.dets.size() == 2 ? r.dets[1] : GQ_NO_DET);
                B->comp_obs.push_back(r.obs);
            } else {
                dc.have = false; dc.omask = r.obs; dc.acc.clear();
                dc.rec(r.dets);
                if (!dc.have) {
                    std::string msg = "Failed to decompose an error into graphlike components:";
                    for (uint32_t d : r.dets) msg += " D" + std::to_string(d);
                    for (int k = 0; k < 64; ++k) if ((r.obs >> k) & 1) msg += " L" + std::to_string(k);
                    delete B;
                    GQ_FAIL(GQ_E_UNSUPPORTED, msg);
                }
                B->comp_a.insert(B->comp_a.end(), dc.best_a.begin(), dc.best_a.end());
                B->comp_b.insert(B->comp_b.end(), dc.best_b.begin(), dc.best_b.end());
                B->comp_obs.insert(B->comp_obs.end(), dc.best_o.begin(), dc.best_o.end());
            }
            B->comp_indptr.push_back((uint32_t)B->comp_a.size());
        }
    }
    *out = B;
    return GQ_OK;
}

extern "C" int gq_dem_build_sizes(const gq_dem_build *b, uint32_t sizes[6]) {
    if (!b || !sizes) GQ_FAIL(GQ_E_INVALID, "gq_dem_build_sizes: NULL argument");
    sizes[0] = b->D; sizes[1] = b->O; sizes[2] = (uint32_t)b->p.size();
    sizes[3] = (uint32_t)b->det_idx.size(); sizes[4] = (uint32_t)b->obs_idx.size(); sizes[5] = (uint32_t)b->comp_a.size();
    return GQ_OK;
}

extern "C" int gq_dem_build_export(const gq_dem_build *b, double *p, uint32_t *det_indptr, uint32_t *det_idx,
                                   uint32_t *obs_indptr, uint32_t *obs_idx, uint32_t *comp_indptr,
                                   uint32_t *comp_a, uint32_t *comp_b, uint64_t *comp_obs) {
    if (!b || !p || !det_indptr || !det_idx || !obs_indptr || !obs_idx) GQ_FAIL(GQ_E_INVALID, "gq_dem_build_export: NULL argument");
    auto cp = [](void *dst, const void *src, size_t n) { if (n) memcpy(dst, src, n); };
    cp(p, b->p.data(), b->p.size() * sizeof(double));
    cp(det_indptr, b->det_indptr.data(), b->det_indptr.size() * 4);
    cp(det_idx, b->det_idx.data(), b->det_idx.size() * 4);
    cp(obs_indptr, b->obs_indptr.data(), b->obs_indptr.size() * 4);
    cp(obs_idx, b->obs_idx.data(), b->obs_idx.size() * 4);
    if (b->decomposed) {
        if (!comp_indptr || !comp_a || !comp_b || !comp_obs) GQ_FAIL(GQ_E_INVALID, "gq_dem_build_export: the model is decomposed, component arrays are required");
        cp(comp_indptr, b->comp_indptr.data(), b->comp_indptr.size() * 4);
        cp(comp_a, b->comp_a.data(), b->comp_a.size() * 4);
        cp(comp_b, b->comp_b.data(), b->comp_b.size() * 4);
        cp(comp_obs, b->comp_obs.data(), b->comp_obs.size() * 8);
    }
    return GQ_OK;
}

extern "C" void gq_dem_build_destroy(gq_dem_build *b) { delete b; }
