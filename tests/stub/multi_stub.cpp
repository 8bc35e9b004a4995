// multi_stub.cpp — CPU harness for the host logic of neopz_b200/csrc/multi.cpp (partition by smallest destination equation,
// local numberings and patterns, staging segments, push maps).  It compiles multi.cpp together with a FAKE b200asm_ctx that
// records what the multi layer hands to every GPU and "assembles" synthetic element matrices on the CPU:
//     ek(i, j) = w(dest_i, dest_j)  with a symmetric integer-valued w, ef(i) = v(dest_i)
// so that sums are exact in double and the global result is known in closed form from the global arrays alone.
// stub_check() runs: local assembly of every fake context -> push along the recorded maps -> comparison of every owned block
// (download window) with the directly assembled global matrix.  Test infrastructure only (tests/test_multi_partition.py);
// the product never links it.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200asm.h"

struct FakeGroup { int m; int64_t nel; std::vector<int64_t> dest; };
struct FakeLink { bool push; b200asm_ctx *peer; int slot_there; int64_t n_a, a_src0; std::vector<int32_t> a_dst, rhs_src, rhs_dst; };

struct b200asm_ctx {
    int device = 0;
    std::vector<FakeGroup> groups;
    int64_t neq = 0, nnz = 0;
    int symmetric = 1;
    std::vector<int64_t> ia, ja;
    std::vector<double> a, rhs;
    std::vector<FakeLink> links;
    int64_t staging_lo = 0, staging_hi = 0, dl_a = -1, dl_r0 = 0, dl_rn = -1;
    int64_t touched_outside_staging = 0;  // entries in non-owned rows that are not staging rows (must stay 0)
    std::string err;
};

static std::string g_err;

extern "C" {
int b200asm_create(b200asm_ctx **out, int device) { *out = new b200asm_ctx(); (*out)->device = device; return 0; }
void b200asm_destroy(b200asm_ctx *c) { delete c; }
const char *b200asm_last_error(const b200asm_ctx *c) { return c ? c->err.c_str() : g_err.c_str(); }
int b200asm_set_option(b200asm_ctx *c, const char *name, int64_t v) {
    const std::string n(name);
    if (n == "staging_lo") c->staging_lo = v;
    else if (n == "staging_hi") c->staging_hi = v;
    else if (n == "download_a_count") c->dl_a = v;
    else if (n == "download_rhs_first") c->dl_r0 = v;
    else if (n == "download_rhs_count") c->dl_rn = v;
    return 0;
}
int b200asm_set_nodes(b200asm_ctx *, int64_t, const double *) { return 0; }
int b200asm_add_group(b200asm_ctx *c, const b200asm_group *g) {
    FakeGroup f;
    f.m = g->nshape * g->nstate;
    f.nel = g->nel;
    f.dest.assign(g->dest, g->dest + (size_t)g->nel * f.m);
    c->groups.push_back(f);
    return (int)c->groups.size() - 1;
}
int b200asm_set_group_coef(b200asm_ctx *, int, const double *) { return 0; }
int b200asm_set_group_force(b200asm_ctx *, int, const double *) { return 0; }
int b200asm_clear_groups(b200asm_ctx *c) { c->groups.clear(); return 0; }
int b200asm_set_pattern(b200asm_ctx *c, int64_t neq, const int64_t *ia, const int64_t *ja, int symmetric) {
    c->neq = neq; c->nnz = ia[neq]; c->symmetric = symmetric;
    c->ia.assign(ia, ia + neq + 1);
    c->ja.assign(ja, ja + c->nnz);
    for (int64_t r = 0; r < neq; r++)
        for (int64_t q = ia[r] + 1; q < ia[r + 1]; q++)
            if (ja[q] <= ja[q - 1]) { c->err = "local row not strictly ascending"; return B200ASM_EINVAL; }
    return 0;
}
int b200asm_exchange_add_peer(b200asm_ctx *c, int push, int slot_there, const b200asm_ipc_mem *, b200asm_ctx *peer, int64_t) {
    c->links.push_back(FakeLink{push != 0, peer, slot_there, 0, 0, {}, {}, {}});
    return (int)c->links.size() - 1;
}
int b200asm_exchange_set_map(b200asm_ctx *c, int link, int64_t n_a, int64_t a_src0, const int32_t *a_dst, int64_t n_rhs, const int32_t *rhs_src,
                             const int32_t *rhs_dst) {
    FakeLink &l = c->links[link];
    l.n_a = n_a; l.a_src0 = a_src0;
    l.a_dst.assign(a_dst, a_dst + n_a);
    l.rhs_src.assign(rhs_src, rhs_src + n_rhs);
    l.rhs_dst.assign(rhs_dst, rhs_dst + n_rhs);
    return 0;
}
int b200asm_exchange_clear(b200asm_ctx *c) { c->links.clear(); return 0; }
int b200asm_synchronize(b200asm_ctx *) { return 0; }
int b200asm_assemble(b200asm_ctx *, double *, double *) { return B200ASM_ENODEVICE; }
int b200asm_assemble_rhs(b200asm_ctx *, double *) { return B200ASM_ENODEVICE; }
int b200asm_assemble_async(b200asm_ctx *) { return B200ASM_ENODEVICE; }
int b200asm_counters(const b200asm_ctx *, int64_t *a, int64_t *b, int64_t *c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return 0; }
}

#include "../../neopz_b200/csrc/cg_sharded.h"
// the solver needs the device: the CPU harness only checks the partition / push maps / halo maps
int b200asm_cg_sharded(int, const b200asm_cg_shard *, int, int64_t, double, int, const double *, double *, int64_t *, double *, std::string &err) {
    err = "stub: no device";
    return B200ASM_ENODEVICE;
}
#include "../../neopz_b200/csrc/multi.cpp"

static double w_of(int64_t a, int64_t b) {  // symmetric, integer valued
    const int64_t lo = std::min(a, b), hi = std::max(a, b);
    return (double)(1 + (lo * 31 + hi * 17) % 97);
}
static double v_of(int64_t a) { return (double)(1 + (a * 13) % 29); }

static int64_t find_pos(const std::vector<int64_t> &ia, const std::vector<int64_t> &ja, int64_t row, int64_t col) {
    const int64_t *b = ja.data() + ia[row], *e = ja.data() + ia[row + 1];
    const int64_t *it = std::lower_bound(b, e, col);
    return (it == e || *it != col) ? -1 : (int64_t)(it - ja.data());
}

// returns 0 when every owned block equals the directly assembled global system; otherwise a positive code; `detail` [8] receives
// counters (missing slots, value mismatches, rhs mismatches, entries outside staging, links, staged entries, ...)
extern "C" int stub_check(b200asm_multi *m, int64_t neq, const int64_t *ia, const int64_t *ja, int symmetric, int64_t *detail) {
    memset(detail, 0, 8 * sizeof(int64_t));
    if (!m->have_pattern) return 1;
    // ---- the global system, assembled directly from the host copies of the groups (global destination indices)
    std::vector<int64_t> gia(ia, ia + neq + 1), gja(ja, ja + ia[neq]);
    std::vector<double> A((size_t)ia[neq], 0.0), F((size_t)neq, 0.0);
    for (const HostGroup &g : m->groups)
        for (int64_t e = 0; e < g.meta.nel; e++) {
            const int64_t *d = &g.dest[(size_t)e * g.m];
            for (int i = 0; i < g.m; i++) {
                if (d[i] < 0) continue;
                F[d[i]] += v_of(d[i]);
                for (int j = symmetric ? i : 0; j < g.m; j++) {
                    if (d[j] < 0) continue;
                    const int64_t row = symmetric ? std::min(d[i], d[j]) : d[i], col = symmetric ? std::max(d[i], d[j]) : d[j];
                    const int64_t pos = find_pos(gia, gja, row, col);
                    if (pos < 0) return 2;
                    A[pos] += w_of(d[i], d[j]);
                }
            }
        }
    // ---- every fake context: local assembly (the values depend on the GLOBAL equation numbers: map back through eqs)
    for (Part &p : m->parts) {
        b200asm_ctx *c = p.ctx;
        c->a.assign((size_t)c->nnz, 0.0);
        c->rhs.assign((size_t)c->neq, 0.0);
        if (c->neq != p.nlocal) return 3;
        const int64_t own0 = p.own_first, own1 = p.own_first + (p.row1 - p.row0);
        for (const FakeGroup &g : c->groups)
            for (int64_t e = 0; e < g.nel; e++) {
                const int64_t *d = &g.dest[(size_t)e * g.m];
                for (int i = 0; i < g.m; i++) {
                    if (d[i] < 0) continue;
                    if (d[i] >= c->neq) return 4;
                    c->rhs[d[i]] += v_of(p.eqs[d[i]]);
                    for (int j = symmetric ? i : 0; j < g.m; j++) {
                        if (d[j] < 0) continue;
                        const int64_t row = symmetric ? std::min(d[i], d[j]) : d[i], col = symmetric ? std::max(d[i], d[j]) : d[j];
                        const int64_t pos = find_pos(c->ia, c->ja, row, col);
                        if (pos < 0) { detail[0]++; continue; }
                        if (!(row >= own0 && row < own1) && !(row >= c->staging_lo && row < c->staging_hi)) detail[3]++;
                        c->a[pos] += w_of(p.eqs[d[i]], p.eqs[d[j]]);
                    }
                }
            }
    }
    // ---- push along the recorded maps
    for (Part &p : m->parts) {
        b200asm_ctx *c = p.ctx;
        std::vector<uint8_t> sent((size_t)c->nnz, 0);
        for (const FakeLink &l : c->links) {
            if (!l.push) continue;
            detail[4]++;
            detail[5] += l.n_a;
            for (int64_t k = 0; k < l.n_a; k++) {
                if (l.a_dst[k] < 0 || l.a_dst[k] >= l.peer->nnz) return 5;
                l.peer->a[l.a_dst[k]] += c->a[l.a_src0 + k];
                sent[l.a_src0 + k] = 1;
            }
            for (size_t k = 0; k < l.rhs_src.size(); k++) {
                if (l.rhs_dst[k] < 0 || l.rhs_dst[k] >= l.peer->neq) return 6;
                l.peer->rhs[l.rhs_dst[k]] += c->rhs[l.rhs_src[k]];
            }
            // the link on the other side must point back
            if (l.slot_there < 0 || l.slot_there >= (int)l.peer->links.size() || l.peer->links[l.slot_there].peer != c || l.peer->links[l.slot_there].push)
                return 7;
        }
        // every staged value travelled
        const int64_t nown = c->dl_a < 0 ? c->nnz : c->dl_a;
        for (int64_t q = nown; q < c->nnz; q++)
            if (c->a[q] != 0.0 && !sent[q]) detail[6]++;
    }
    // ---- owned blocks against the global system
    for (Part &p : m->parts) {
        b200asm_ctx *c = p.ctx;
        const int64_t nown = c->dl_a < 0 ? c->nnz : c->dl_a;
        if (nown != gia[p.row1] - gia[p.row0]) return 8;
        for (int64_t q = 0; q < nown; q++)
            if (c->a[q] != A[gia[p.row0] + q]) detail[1]++;
        if (c->dl_rn != p.row1 - p.row0) return 9;
        for (int64_t r = 0; r < c->dl_rn; r++)
            if (c->rhs[c->dl_r0 + r] != F[p.row0 + r]) detail[2]++;
    }
    return (detail[0] || detail[1] || detail[2] || detail[3] || detail[6]) ? 10 : 0;
}
