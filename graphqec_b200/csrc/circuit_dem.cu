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

const uint32_t SEP = 0xFFFFFFFFu;   // component separator inside a decomposed key ("^" in Stim's text)

inline int sym_ndet(const Sym &s) {
    int n = 0;
    for (uint32_t x : s) n += !(x & OBS_BIT);
    return n;
}

// decompose_errors = True, restated after Stim's ErrorAnalyzer; the two stages are described in
// graphqec_b200/dem.py (the readable statement of the same algorithm; tests/test_dem.py keeps the two bit-identical).
//
// Stage 1: keys (component lists) of the Pauli products k = 1 .. n - 1 of ONE channel.  sym[k] = symptom of product k.
// Returns false when the channel touches more than 15 detectors (Stim: "too many detectors to find reducible errors").
bool in_channel_keys(const Sym *sym, int n, std::vector<Sym> &keys, std::string &err) {
    keys.assign(n, Sym());
    int cnt[16], maxc = 0;
    for (int k = 1; k < n; ++k) { keys[k] = sym[k]; cnt[k] = sym_ndet(sym[k]); maxc = std::max(maxc, cnt[k]); }
    if (maxc <= 2) return true;
    Sym involved;
    for (int k = 1; k < n; ++k) for (uint32_t x : sym[k]) if (!(x & OBS_BIT)) involved.push_back(x);
    std::sort(involved.begin(), involved.end());
    involved.erase(std::unique(involved.begin(), involved.end()), involved.end());
    if (involved.size() > 15) { err = "An error involves too many detectors (>15) to find reducible errors."; return false; }
    uint32_t dm[16];
    for (int k = 1; k < n; ++k) {
        dm[k] = 0;
        for (uint32_t x : sym[k]) if (!(x & OBS_BIT)) dm[k] |= 1u << (std::lower_bound(involved.begin(), involved.end(), x) - involved.begin());
    }
    uint32_t singles = 0;
    for (int k = 1; k < n; ++k) if (cnt[k] == 1) singles |= dm[k];
    int irr[16], nirr = 0;
    for (int k = 1; k < n; ++k) if (cnt[k] == 2 && (dm[k] & ~singles)) irr[nirr++] = k;
    Sym acc, tmp;
    for (int g = 1; g < n; ++g) {
        if (cnt[g] <= 2) continue;
        const uint32_t goal = dm[g];
        int parts[16], np = 0;
        uint32_t rem = goal;
        if (goal & ~singles) {
            bool found = false;
            for (int i = 0; i < nirr && !found; ++i) {
                const uint32_t mk = dm[irr[i]];
                if ((goal & mk) == mk && (goal & ~(singles | mk)) == 0) { parts[0] = irr[i]; np = 1; rem = goal & ~mk; found = true; }
            }
            for (int i1 = 0; i1 < nirr && !found; ++i1) {
                const uint32_t m1 = dm[irr[i1]];
                if ((goal & m1) != m1) continue;
                for (int i2 = i1 + 1; i2 < nirr && !found; ++i2) {
                    const uint32_t m2 = dm[irr[i2]];
                    if ((m1 & m2) == 0 && (goal & m2) == m2 && (goal & ~(singles | m1 | m2)) == 0) {
                        parts[0] = irr[i1]; parts[1] = irr[i2]; np = 2; rem = goal & ~(m1 | m2); found = true;
                    }
                }
            }
            if (!found) continue;
        }
        for (int k = 1; k < n && rem; ++k)
            if (cnt[k] == 1 && (dm[k] & ~rem) == 0) { rem &= ~dm[k]; parts[np++] = k; }
        if (rem) continue;
        acc.clear();
        for (int i = 0; i < np; ++i) sym_xor(acc, sym[parts[i]], tmp);
        if (acc != sym[g]) continue;   // the pieces must reproduce the observables too, else stage 2 deals with it
        Sym &key = keys[g];
        key.clear();
        for (int i = 0; i < np; ++i) {
            if (i) key.push_back(SEP);
            key.insert(key.end(), sym[parts[i]].begin(), sym[parts[i]].end());
        }
    }
    return true;
}

struct Entry { Sym key; double p; };

// Stage 2: greedy global pass over all errors in Stim order (plain lexicographic order of the keys: detector ids <
// OBS_BIT | observable < SEP); errors that end up with the same decomposition merge.  false = decomposition failure.
bool global_decomposition_pass(std::vector<Entry> &entries, std::string &err) {
    std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.key < b.key; });
    std::map<std::pair<uint32_t, uint32_t>, Sym> known;   // detectors of a graphlike component -> the first such component
    auto each_component = [](const Sym &key, auto &&fn) {
        size_t b = 0;
        for (size_t i = 0; i <= key.size(); ++i)
            if (i == key.size() || key[i] == SEP) { fn(key.data() + b, key.data() + i); b = i + 1; }
    };
    for (const Entry &e : entries)
        each_component(e.key, [&](const uint32_t *b, const uint32_t *en) {
            uint32_t d[3]; int nd = 0;
            for (const uint32_t *q = b; q < en; ++q) if (!(*q & OBS_BIT)) { if (nd < 3) d[nd] = *q; ++nd; }
            if (nd == 1 || nd == 2) known.emplace(std::make_pair(d[0], nd == 2 ? d[1] : GQ_NO_DET), Sym(b, en));
        });
    std::vector<Entry> out;
    std::unordered_map<Sym, uint32_t, SymHash> index;
    Sym left, tmp, dets;
    for (const Entry &e : entries) {
        Sym nk;
        bool ok = true;
        auto push = [&](const Sym &c) { if (!nk.empty()) nk.push_back(SEP); nk.insert(nk.end(), c.begin(), c.end()); };
        each_component(e.key, [&](const uint32_t *b, const uint32_t *en) {
            if (!ok) return;
            dets.clear();
            for (const uint32_t *q = b; q < en; ++q) if (!(*q & OBS_BIT)) dets.push_back(*q);
            if (dets.size() <= 2) { push(Sym(b, en)); return; }
            std::vector<char> done(dets.size(), 0);
            left.assign(b, en);
            for (size_t i = 0; i < dets.size(); ++i) {
                if (done[i]) continue;
                for (size_t j = i + 1; j < dets.size(); ++j) {
                    if (done[j]) continue;
                    auto it = known.find(std::make_pair(dets[i], dets[j]));
                    if (it != known.end()) { done[i] = done[j] = 1; push(it->second); sym_xor(left, it->second, tmp); break; }
                }
            }
            int missed = 0;
            for (size_t i = 0; i < dets.size(); ++i) {
                if (done[i]) continue;
                auto it = known.find(std::make_pair(dets[i], GQ_NO_DET));
                if (it != known.end()) { done[i] = 1; push(it->second); sym_xor(left, it->second, tmp); }
                else ++missed;
            }
            if (missed > 2 || (!left.empty() && sym_ndet(left) == 0)) {
                err = "Failed to decompose errors into graphlike components with at most two symptoms. The error component that failed to decompose is '";
                bool first = true;
                for (const uint32_t *q = b; q < en; ++q) {
                    if (!first) err += ", ";
                    first = false;
                    err += (*q & OBS_BIT) ? "L" + std::to_string(*q & 63) : "D" + std::to_string(*q);
                }
                err += "'.";
                ok = false;
                return;
            }
            if (!left.empty()) push(left);   // remnant edge
        });
        if (!ok) return false;
        auto it = index.find(nk);
        if (it == index.end()) { index.emplace(nk, (uint32_t)out.size()); out.push_back(Entry{nk, e.p}); }
        else { double &old = out[it->second].p; old = old * (1.0 - e.p) + e.p * (1.0 - old); }
    }
    entries.swap(out);
    return true;
}

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
    // one noise channel instance: s basis Paulis -> 2^s - 1 products, each an independent error of probability pr
    Sym csym[16];
    std::vector<Sym> ckeys;
    std::string cerr;
    bool channel_failed = false;
    auto channel = [&](const Sym *const *basis, int s, double pr) {
        if (pr <= 0.0 || channel_failed) return;
        const int n = 1 << s;
        for (int k = 1; k < n; ++k) {
            const int low = k & -k, b = __builtin_ctz((unsigned)k);
            csym[k] = csym[k ^ low];
            sym_xor(csym[k], *basis[b], tmp);
        }
        if (!decompose_errors) {
            for (int k = 1; k < n; ++k) emit(csym[k], pr);
            return;
        }
        if (!in_channel_keys(csym, n, ckeys, cerr)) { channel_failed = true; return; }
        for (int k = 1; k < n; ++k) if (!csym[k].empty()) emit(ckeys[k], pr);
    };
    csym[0].clear();
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
                const Sym *basis[2] = {&xs[q], &zs[q]};
                channel(basis, 2, q1);
            }
            break;
        }
        case G_XERR: for (uint32_t k = op.t0; k < op.t1; ++k) { const Sym *basis[1] = {&xs[t[k]]}; channel(basis, 1, op.has_arg ? op.arg : 0.0); } break;
        case G_ZERR: for (uint32_t k = op.t0; k < op.t1; ++k) { const Sym *basis[1] = {&zs[t[k]]}; channel(basis, 1, op.has_arg ? op.arg : 0.0); } break;
        case G_YERR:
            for (uint32_t k = op.t0; k < op.t1; ++k) {
                acc2 = xs[t[k]]; sym_xor(acc2, zs[t[k]], tmp);
                const Sym *basis[1] = {&acc2};
                channel(basis, 1, op.has_arg ? op.arg : 0.0);
            }
            break;
        case G_DEP2: {
            const double pr = op.has_arg ? op.arg : 0.0;
            if (pr > 15.0 / 16.0) GQ_FAIL(GQ_E_INVALID, "DEPOLARIZE2 probability above 15/16 has no independent-component form");
            const double q2 = 0.5 - 0.5 * pow(1.0 - 16.0 * pr / 15.0, 0.125);
            for (uint32_t k = op.t0; k < op.t1; k += 2) {
                const Sym *basis[4] = {&xs[t[k]], &zs[t[k]], &xs[t[k + 1]], &zs[t[k + 1]]};
                channel(basis, 4, q2);
            }
            break;
        }
        case G_M: case G_MX:
            for (uint32_t k = op.t1; k-- > op.t0;) {
                --m;
                sym_xor(op.gate == G_M ? xs[t[k]] : zs[t[k]], mmask[m], tmp);
                { const Sym *basis[1] = {&mmask[m]}; channel(basis, 1, op.has_arg ? op.arg : 0.0); }
            }
            break;
        case G_MR: case G_MRX:
            for (uint32_t k = op.t1; k-- > op.t0;) {
                --m;
                if (op.gate == G_MR) { xs[t[k]] = mmask[m]; zs[t[k]].clear(); }
                else { zs[t[k]] = mmask[m]; xs[t[k]].clear(); }
                { const Sym *basis[1] = {&mmask[m]}; channel(basis, 1, op.has_arg ? op.arg : 0.0); }
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

    if (channel_failed) GQ_FAIL(GQ_E_UNSUPPORTED, cerr);
    gq_dem_build *B = new gq_dem_build();
    B->D = ndet; B->O = nobs;
    B->det_indptr.push_back(0); B->obs_indptr.push_back(0); B->comp_indptr.push_back(0);
    B->decomposed = decompose_errors != 0;
    auto push_row = [&](const Sym &whole, double pr) {
        B->p.push_back(pr);
        for (uint32_t id : whole) {
            if (id & OBS_BIT) B->obs_idx.push_back(id & 63);
            else B->det_idx.push_back(id);
        }
        B->det_indptr.push_back((uint32_t)B->det_idx.size());
        B->obs_indptr.push_back((uint32_t)B->obs_idx.size());
    };
    if (!decompose_errors) {
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
        for (const Row &r : rows) {
            Sym whole(r.dets);
            for (int k = 0; k < 64; ++k) if ((r.obs >> k) & 1) whole.push_back(OBS_BIT | (uint32_t)k);
            push_row(whole, r.p);
        }
    } else {
        // ---- Stim's second stage, then one row per distinct decomposition, in Stim's order
        std::vector<Entry> entries(probs.size());
        for (size_t i = 0; i < probs.size(); ++i) { entries[i].key = *keys[i]; entries[i].p = probs[i]; }
        std::string derr;
        if (!global_decomposition_pass(entries, derr)) { delete B; GQ_FAIL(GQ_E_UNSUPPORTED, derr); }
        std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.key < b.key; });
        Sym whole, comp;
        for (const Entry &e : entries) {
            whole.clear();
            size_t b0 = 0;
            for (size_t i = 0; i <= e.key.size(); ++i) {
                if (i < e.key.size() && e.key[i] != SEP) continue;
                comp.assign(e.key.begin() + b0, e.key.begin() + i);
                b0 = i + 1;
                sym_xor(whole, comp, tmp);
                uint32_t d[2] = {GQ_NO_DET, GQ_NO_DET};
                int nd = 0;
                uint64_t om = 0;
                for (uint32_t id : comp) {
                    if (id & OBS_BIT) om |= 1ull << (id & 63);
                    else if (nd < 2) d[nd++] = id;
                }
                B->comp_a.push_back(d[0]); B->comp_b.push_back(d[1]); B->comp_obs.push_back(om);
            }
            B->comp_indptr.push_back((uint32_t)B->comp_a.size());
            push_row(whole, e.p);
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
