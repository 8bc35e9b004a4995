// multi.cpp — b200asm_multi: one assembly spread over several GPUs of one process (include/b200asm.h, "multi-GPU in one
// process").  It is the GPU counterpart of the reference's only parallel knob, the thread count of the strategy
// (StrMatrix/TPZStrMatParInterface.h:62-69, dispatched at StrMatrix/pzstrmatrixor.cpp:64-68,476-513): the caller hands over the
// SAME flattened mesh and the SAME global CSR pattern it would give one context; this layer
//   1. partitions the elements by their smallest destination equation into contiguous chunks of equal work: GPU g
//      then owns the row block [r_g, r_g+1) of the global CSR, and every entry of an element of GPU g lies in a row >= r_g;
//   2. builds, per GPU, a local system in a numbering that is monotone in the global one: the owned rows with their complete
//      global pattern, followed by "staging" rows = the entries its elements contribute to rows of other GPUs (exactly those
//      entries, found from the interface elements), and hands it to an ordinary b200asm_ctx;
//   3. links the contexts (b200asm_exchange_*): staged values are added into the owner's memory over NVLink while the interior
//      elements are assembled; the owned blocks are then copied to the caller's global arrays, every GPU its own slice.
// Pure host code over the public C ABI: no kernels here.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200asm.h"
#include "cg_sharded.h"

namespace {

struct HostGroup {
    b200asm_group meta{};
    std::vector<int32_t> elnodes;
    std::vector<int64_t> dest;
    std::vector<double> qpts, qwts, phi, dphi, force;
    int ncorner = 0, m = 0, dim = 3;
};

int ncorner_of(int topology) {
    switch (topology) {
        case B200ASM_HEX: return 8;
        case B200ASM_TET: return 4;
        case B200ASM_QUAD: return 4;
        case B200ASM_TRI: return 3;
        case B200ASM_LINE: return 2;
        case B200ASM_PRISM: return 6;
        case B200ASM_PYRAMID: return 5;
    }
    return -1;
}
int dim_of(int topology) {
    return (topology == B200ASM_QUAD || topology == B200ASM_TRI) ? 2 : (topology == B200ASM_LINE ? 1 : 3);
}

struct Part {
    b200asm_ctx *ctx = nullptr;
    int device = 0;
    int64_t row0 = 0, row1 = 0;        // owned global rows
    int64_t own_first = 0;             // local index of global row row0
    int64_t nlocal = 0, nnz_owned = 0, nnz_local = 0;
    std::vector<int64_t> eqs;          // local -> global equation (ascending)
    std::vector<std::vector<int64_t>> elements;  // per global group: element indices assigned to this GPU
    std::vector<int> local_group;      // per global group: index of its sub-group in ctx (-1: none)
    int64_t nelements = 0;
    // halo of the owned rows (sharded CG): local equations owned by other GPUs that appear as columns of the owned rows
    std::vector<int32_t> halo_local, halo_owner, halo_remote;
};

}  // namespace

struct b200asm_multi {
    std::vector<Part> parts;
    std::vector<HostGroup> groups;
    std::vector<double> xyz;
    int64_t nnodes = 0;
    int64_t neq = 0, nnz = 0;
    int symmetric = 1;
    bool have_pattern = false;
    std::vector<int64_t> ia;           // global row pointers (kept: slices of the caller's arrays)
    std::vector<std::pair<std::string, int64_t>> options;
    std::string err;
};

namespace {

thread_local std::string g_multi_create_error;

int mfail(b200asm_multi *m, int code, const std::string &msg) {
    if (m) m->err = msg; else g_multi_create_error = msg;
    return code;
}
int pass(b200asm_multi *m, b200asm_ctx *ctx, int rc, const char *what) {
    if (rc < 0) return mfail(m, rc, std::string(what) + ": " + b200asm_last_error(ctx));
    return rc;
}

// smallest destination equation of an element (-1: every equation filtered)
int64_t min_dest(const int64_t *d, int m) {
    int64_t mn = INT64_MAX;
    for (int k = 0; k < m; k++)
        if (d[k] >= 0 && d[k] < mn) mn = d[k];
    return mn == INT64_MAX ? -1 : mn;
}

}  // namespace

extern "C" const char *b200asm_multi_last_error(const b200asm_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

extern "C" int b200asm_multi_create(b200asm_multi **out, int ndev, const int *devices) {
    if (!out || ndev < 1 || ndev > B200ASM_MAX_PEERS / 2) return mfail(nullptr, B200ASM_EINVAL, "b200asm_multi_create: 1..8 devices");
    *out = nullptr;
    std::unique_ptr<b200asm_multi> m(new b200asm_multi());
    m->parts.resize(ndev);
    for (int k = 0; k < ndev; k++) {
        Part &p = m->parts[k];
        p.device = devices ? devices[k] : k;
        for (int j = 0; j < k; j++)
            if (m->parts[j].device == p.device) {
                for (int i = 0; i < k; i++) b200asm_destroy(m->parts[i].ctx);
                return mfail(nullptr, B200ASM_EINVAL, "b200asm_multi_create: a device is listed twice");
            }
        const int rc = b200asm_create(&p.ctx, p.device);
        if (rc != 0) {
            const std::string msg = b200asm_last_error(nullptr);
            for (int j = 0; j < k; j++) b200asm_destroy(m->parts[j].ctx);
            return mfail(nullptr, rc, "b200asm_multi_create: device " + std::to_string(p.device) + ": " + msg);
        }
    }
    *out = m.release();
    return 0;
}

extern "C" void b200asm_multi_destroy(b200asm_multi *m) {
    if (!m) return;
    // links first: no context may be freed while another still maps its arrays
    for (Part &p : m->parts) b200asm_synchronize(p.ctx);
    for (Part &p : m->parts) b200asm_exchange_clear(p.ctx);
    for (Part &p : m->parts) b200asm_destroy(p.ctx);
    delete m;
}

extern "C" int b200asm_multi_num_devices(const b200asm_multi *m) { return m ? (int)m->parts.size() : B200ASM_EINVAL; }

extern "C" int b200asm_multi_context(b200asm_multi *m, int k, b200asm_ctx **ctx) {
    if (!m || !ctx || k < 0 || k >= (int)m->parts.size()) return mfail(m, B200ASM_EINVAL, "b200asm_multi_context: bad arguments");
    *ctx = m->parts[k].ctx;
    return 0;
}

extern "C" int b200asm_multi_set_option(b200asm_multi *m, const char *name, int64_t value) {
    if (!m || !name) return B200ASM_EINVAL;
    for (Part &p : m->parts) {
        const int rc = pass(m, p.ctx, b200asm_set_option(p.ctx, name, value), "b200asm_set_option");
        if (rc < 0) return rc;
    }
    return 0;
}

extern "C" int b200asm_multi_set_nodes(b200asm_multi *m, int64_t nnodes, const double *xyz) {
    if (!m || nnodes < 0 || (nnodes && !xyz)) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_nodes: bad arguments");
    m->xyz.assign(xyz, xyz + (size_t)nnodes * 3);
    m->nnodes = nnodes;
    for (Part &p : m->parts) {  // every GPU holds the whole node table (8 * 3 bytes per node)
        const int rc = pass(m, p.ctx, b200asm_set_nodes(p.ctx, nnodes, xyz), "b200asm_set_nodes");
        if (rc < 0) return rc;
    }
    return 0;
}

extern "C" int b200asm_multi_add_group(b200asm_multi *m, const b200asm_group *gi) {
    if (!m || !gi) return B200ASM_EINVAL;
    HostGroup g;
    g.meta = *gi;
    g.ncorner = ncorner_of(gi->topology);
    g.dim = dim_of(gi->topology);
    if (g.ncorner < 0 || gi->nel < 0 || gi->nshape < 1 || gi->nstate < 1 || gi->nqp < 1 || !gi->elnodes || !gi->dest || !gi->qpts ||
        !gi->qwts || !gi->phi || !gi->dphi)
        return mfail(m, B200ASM_EINVAL, "b200asm_multi_add_group: bad group");
    g.m = gi->nshape * gi->nstate;
    g.elnodes.assign(gi->elnodes, gi->elnodes + (size_t)gi->nel * g.ncorner);
    g.dest.assign(gi->dest, gi->dest + (size_t)gi->nel * g.m);
    g.qpts.assign(gi->qpts, gi->qpts + (size_t)gi->nqp * g.dim);
    g.qwts.assign(gi->qwts, gi->qwts + gi->nqp);
    g.phi.assign(gi->phi, gi->phi + (size_t)gi->nqp * gi->nshape);
    g.dphi.assign(gi->dphi, gi->dphi + (size_t)gi->nqp * g.dim * gi->nshape);
    if (gi->force) g.force.assign(gi->force, gi->force + (size_t)gi->nel * gi->nqp * gi->nstate);
    m->groups.push_back(std::move(g));
    m->have_pattern = false;  // the partition follows the elements: b200asm_multi_set_pattern again
    return (int)m->groups.size() - 1;
}

extern "C" int b200asm_multi_set_group_coef(b200asm_multi *m, int group, const double coef[16]) {
    if (!m || group < 0 || group >= (int)m->groups.size() || !coef) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_group_coef: bad arguments");
    memcpy(m->groups[group].meta.coef, coef, sizeof(double) * 16);
    if (m->have_pattern)
        for (Part &p : m->parts)
            if (p.local_group[group] >= 0) {
                const int rc = pass(m, p.ctx, b200asm_set_group_coef(p.ctx, p.local_group[group], coef), "b200asm_set_group_coef");
                if (rc < 0) return rc;
            }
    return 0;
}

extern "C" int b200asm_multi_set_group_force(b200asm_multi *m, int group, const double *force) {
    if (!m || group < 0 || group >= (int)m->groups.size()) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_group_force: bad arguments");
    HostGroup &g = m->groups[group];
    const size_t per = (size_t)g.meta.nqp * g.meta.nstate;
    if ((force != nullptr) != !g.force.empty()) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_group_force: the group was added with / without a table");
    if (!force) return 0;
    g.force.assign(force, force + (size_t)g.meta.nel * per);
    if (m->have_pattern)
        for (Part &p : m->parts)
            if (p.local_group[group] >= 0) {
                const std::vector<int64_t> &el = p.elements[group];
                std::vector<double> sub(el.size() * per);
                for (size_t k = 0; k < el.size(); k++) memcpy(&sub[k * per], &g.force[(size_t)el[k] * per], per * sizeof(double));
                const int rc = pass(m, p.ctx, b200asm_set_group_force(p.ctx, p.local_group[group], sub.data()), "b200asm_set_group_force");
                if (rc < 0) return rc;
            }
    return 0;
}

extern "C" int b200asm_multi_clear_groups(b200asm_multi *m) {
    if (!m) return B200ASM_EINVAL;
    m->groups.clear();
    m->have_pattern = false;
    for (Part &p : m->parts) {
        b200asm_exchange_clear(p.ctx);
        b200asm_clear_groups(p.ctx);
        p.elements.clear();
        p.local_group.clear();
    }
    return 0;
}

extern "C" int b200asm_multi_set_pattern(b200asm_multi *m, int64_t neq, const int64_t *ia, const int64_t *ja, int symmetric) {
    if (!m || neq < 0 || !ia || (ia[neq] && !ja)) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: bad arguments");
    const int ndev = (int)m->parts.size();
    const size_t ng = m->groups.size();
    m->have_pattern = false;
    m->neq = neq; m->nnz = ia[neq]; m->symmetric = symmetric ? 1 : 0;
    m->ia.assign(ia, ia + neq + 1);
    for (Part &p : m->parts) {
        b200asm_synchronize(p.ctx);
        b200asm_exchange_clear(p.ctx);
        b200asm_clear_groups(p.ctx);
    }

    // ---- 1. partition: elements sorted by their smallest destination equation, cut into chunks of equal work ----------
    struct El { int64_t mind; int32_t group; int64_t e; double work; };
    std::vector<El> els;
    {
        size_t total = 0;
        for (const HostGroup &g : m->groups) total += (size_t)g.meta.nel;
        els.reserve(total);
    }
    for (size_t gi = 0; gi < ng; gi++) {
        const HostGroup &g = m->groups[gi];
        // cost model: the Gram products (upper triangle of ndof^2 per point) dominate
        const double w = (double)g.m * g.m * g.meta.nqp + 64.0;
        for (int64_t e = 0; e < g.meta.nel; e++) {
            const int64_t mn = min_dest(&g.dest[(size_t)e * g.m], g.m);
            for (int k = 0; k < g.m; k++)
                if (g.dest[(size_t)e * g.m + k] >= neq) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: a destination index exceeds the pattern");
            els.push_back({mn < 0 ? 0 : mn, (int32_t)gi, e, w});
        }
    }
    std::stable_sort(els.begin(), els.end(), [](const El &a, const El &b) { return a.mind < b.mind; });
    double total_work = 0.0;
    for (const El &e : els) total_work += e.work;
    std::vector<int64_t> row_begin(ndev + 1, neq);
    row_begin[0] = 0;
    {
        double acc = 0.0;
        int next = 1;
        for (size_t k = 0; k < els.size() && next < ndev; k++) {
            // a cut falls between two elements with different smallest equations
            if (k > 0 && els[k].mind != els[k - 1].mind && acc >= total_work * next / ndev) {
                row_begin[next] = els[k].mind;
                next++;
            }
            acc += els[k].work;
        }
        for (int k = 1; k <= ndev; k++) row_begin[k] = std::max(row_begin[k], row_begin[k - 1]);
    }
    // owner of a row: the last GPU whose block starts at or below it (a GPU with an empty block owns nothing)
    auto owner_of = [&](int64_t row) {
        return (int)(std::upper_bound(row_begin.begin(), row_begin.begin() + ndev, row) - row_begin.begin()) - 1;
    };
    for (int k = 0; k < ndev; k++) {
        Part &p = m->parts[k];
        p.row0 = row_begin[k]; p.row1 = row_begin[k + 1];
        p.elements.assign(ng, {});
        p.local_group.assign(ng, -1);
        p.nelements = 0;
    }
    for (const El &e : els) {
        Part &p = m->parts[owner_of(e.mind)];
        p.elements[e.group].push_back(e.e);
        p.nelements++;
    }
    for (Part &p : m->parts)
        for (auto &v : p.elements) std::sort(v.begin(), v.end());  // mesh order inside every GPU
    els.clear();
    els.shrink_to_fit();

    // ---- 2. local systems -------------------------------------------------------------------------------------------------
    struct Remote { int64_t row, col; };
    std::vector<std::vector<Remote>> remote(ndev);      // staged entries of every GPU, sorted by (row, col)
    std::vector<std::vector<int64_t>> remote_rows(ndev);  // touched rows above the owned block (load vector)
    std::vector<int32_t> g2l((size_t)neq);
    for (int k = 0; k < ndev; k++) {
        Part &p = m->parts[k];
        std::vector<uint8_t> mark((size_t)neq, 0);  // 1: in the local numbering
        for (int64_t r = p.row0; r < p.row1; r++) mark[r] = 1;
        if (neq) for (int64_t q = ia[p.row0]; q < ia[p.row1]; q++) mark[ja[q]] = 1;
        std::vector<Remote> &rem = remote[k];
        for (size_t gi = 0; gi < ng; gi++) {
            const HostGroup &g = m->groups[gi];
            for (int64_t e : p.elements[gi]) {
                const int64_t *d = &g.dest[(size_t)e * g.m];
                int64_t mx = -1;
                for (int i = 0; i < g.m; i++) {
                    if (d[i] < 0) continue;
                    if (d[i] < p.row0) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: internal: element below its block");
                    mark[d[i]] = 1;
                    mx = std::max(mx, d[i]);
                }
                if (mx < p.row1) continue;
                // interface element: its entries in rows another GPU owns
                for (int i = 0; i < g.m; i++) {
                    if (d[i] < 0) continue;
                    for (int j = symmetric ? i : 0; j < g.m; j++) {
                        if (d[j] < 0) continue;
                        const int64_t row = symmetric ? std::min(d[i], d[j]) : d[i];
                        const int64_t col = symmetric ? std::max(d[i], d[j]) : d[j];
                        if (row >= p.row1) rem.push_back({row, col});
                    }
                }
            }
        }
        std::sort(rem.begin(), rem.end(), [](const Remote &a, const Remote &b) { return a.row != b.row ? a.row < b.row : a.col < b.col; });
        rem.erase(std::unique(rem.begin(), rem.end(), [](const Remote &a, const Remote &b) { return a.row == b.row && a.col == b.col; }), rem.end());
        p.eqs.clear();
        for (int64_t r = 0; r < neq; r++)
            if (mark[r]) {
                g2l[r] = (int32_t)p.eqs.size();
                p.eqs.push_back(r);
            }
        p.nlocal = (int64_t)p.eqs.size();
        p.own_first = p.row1 > p.row0 ? g2l[p.row0] : 0;
        p.nnz_owned = neq ? ia[p.row1] - ia[p.row0] : 0;
        p.nnz_local = p.nnz_owned + (int64_t)rem.size();
        if (p.nnz_local > 0x7fffffff) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: more than 2^31 entries on one GPU: use more devices");
        // local CSR: owned rows with the global pattern, staged entries behind them, nothing in column-only rows
        std::vector<int64_t> lia((size_t)p.nlocal + 1, 0), lja((size_t)p.nnz_local);
        {
            size_t rk = 0;
            int64_t pos = 0;
            std::vector<int64_t> &rrows = remote_rows[k];
            rrows.clear();
            for (int64_t l = 0; l < p.nlocal; l++) {
                const int64_t r = p.eqs[l];
                lia[l] = pos;
                if (r >= p.row0 && r < p.row1) {
                    for (int64_t q = ia[r]; q < ia[r + 1]; q++) {
                        lja[pos++] = g2l[ja[q]];
                        if (ja[q] < p.row0 || ja[q] >= p.row1) mark[ja[q]] = 2;  // a column another GPU owns: halo of the solver
                    }
                } else {
                    while (rk < rem.size() && rem[rk].row < r) rk++;
                    while (rk < rem.size() && rem[rk].row == r) lja[pos++] = g2l[rem[rk++].col];
                }
            }
            lia[p.nlocal] = pos;
            if (pos != p.nnz_local) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: internal: local pattern size");
            p.halo_local.clear();
            for (int64_t l = 0; l < p.nlocal; l++)
                if (mark[p.eqs[l]] == 2) p.halo_local.push_back((int32_t)l);
            // rows above the owned block that an element of this GPU touches (their load-vector entries travel too)
            std::vector<uint8_t> touched((size_t)neq, 0);
            for (size_t gi = 0; gi < ng; gi++) {
                const HostGroup &g = m->groups[gi];
                for (int64_t e : p.elements[gi])
                    for (int i = 0; i < g.m; i++) {
                        const int64_t d = g.dest[(size_t)e * g.m + i];
                        if (d >= p.row1) touched[d] = 1;
                    }
            }
            for (int64_t r = p.row1; r < neq; r++)
                if (touched[r]) rrows.push_back(r);
        }
        // the context: options, staging rows, sub-groups in the local numbering, the local pattern
        int rc;
        const int64_t stage_lo = p.own_first + (p.row1 - p.row0);
        if ((rc = pass(m, p.ctx, b200asm_set_option(p.ctx, "staging_lo", stage_lo), "staging_lo")) < 0) return rc;
        if ((rc = pass(m, p.ctx, b200asm_set_option(p.ctx, "staging_hi", p.nlocal), "staging_hi")) < 0) return rc;
        if ((rc = pass(m, p.ctx, b200asm_set_option(p.ctx, "download_a_count", p.nnz_owned), "download_a_count")) < 0) return rc;
        if ((rc = pass(m, p.ctx, b200asm_set_option(p.ctx, "download_rhs_first", p.own_first), "download_rhs_first")) < 0) return rc;
        if ((rc = pass(m, p.ctx, b200asm_set_option(p.ctx, "download_rhs_count", p.row1 - p.row0), "download_rhs_count")) < 0) return rc;
        for (size_t gi = 0; gi < ng; gi++) {
            const HostGroup &g = m->groups[gi];
            const std::vector<int64_t> &el = p.elements[gi];
            if (el.empty()) continue;
            std::vector<int32_t> en(el.size() * g.ncorner);
            std::vector<int64_t> dl(el.size() * g.m);
            const size_t per = (size_t)g.meta.nqp * g.meta.nstate;
            std::vector<double> fo(g.force.empty() ? 0 : el.size() * per);
            for (size_t k2 = 0; k2 < el.size(); k2++) {
                memcpy(&en[k2 * g.ncorner], &g.elnodes[(size_t)el[k2] * g.ncorner], g.ncorner * sizeof(int32_t));
                for (int i = 0; i < g.m; i++) {
                    const int64_t d = g.dest[(size_t)el[k2] * g.m + i];
                    dl[k2 * g.m + i] = d < 0 ? -1 : g2l[d];
                }
                if (!fo.empty()) memcpy(&fo[k2 * per], &g.force[(size_t)el[k2] * per], per * sizeof(double));
            }
            b200asm_group sub = g.meta;
            sub.nel = (int64_t)el.size();
            sub.elnodes = en.data(); sub.dest = dl.data();
            sub.qpts = g.qpts.data(); sub.qwts = g.qwts.data(); sub.phi = g.phi.data(); sub.dphi = g.dphi.data();
            sub.force = fo.empty() ? nullptr : fo.data();
            if ((rc = pass(m, p.ctx, b200asm_add_group(p.ctx, &sub), "b200asm_add_group")) < 0) return rc;
            p.local_group[gi] = rc;
        }
        if ((rc = pass(m, p.ctx, b200asm_set_pattern(p.ctx, p.nlocal, lia.data(), lja.data(), symmetric), "b200asm_set_pattern")) < 0) return rc;
    }

    // ---- 3. links: where every staged entry goes ----------------------------------------------------------------------------
    std::vector<int> nlinks(ndev, 0);
    for (int k = 0; k < ndev; k++) {
        Part &p = m->parts[k];
        const std::vector<Remote> &rem = remote[k];
        const std::vector<int64_t> &rrows = remote_rows[k];
        size_t a0 = 0, r0 = 0;
        for (int h = k + 1; h < ndev; h++) {
            Part &o = m->parts[h];
            size_t a1 = a0, r1 = r0;
            while (a1 < rem.size() && rem[a1].row < o.row1) a1++;
            while (r1 < rrows.size() && rrows[r1] < o.row1) r1++;
            if (a1 == a0 && r1 == r0) continue;
            std::vector<int32_t> a_dst(a1 - a0), rhs_src(r1 - r0), rhs_dst(r1 - r0);
            for (size_t q = a0; q < a1; q++) {
                const int64_t *b = ja + ia[rem[q].row], *e = ja + ia[rem[q].row + 1];
                const int64_t *it = std::lower_bound(b, e, rem[q].col);
                if (it == e || *it != rem[q].col)
                    return mfail(m, B200ASM_EPATTERN, "b200asm_multi_set_pattern: an element entry has no position in the CSR pattern");
                a_dst[q - a0] = (int32_t)((it - ja) - ia[o.row0]);
            }
            for (size_t q = r0; q < r1; q++) {
                rhs_src[q - r0] = (int32_t)(std::lower_bound(p.eqs.begin(), p.eqs.end(), rrows[q]) - p.eqs.begin());
                rhs_dst[q - r0] = (int32_t)(o.own_first + (rrows[q] - o.row0));
            }
            const int here = nlinks[k]++, there = nlinks[h]++;
            int rc;
            const int64_t in_min = r1 > r0 ? o.own_first + (rrows[r0] - o.row0) : -1;
            if ((rc = pass(m, p.ctx, b200asm_exchange_add_peer(p.ctx, 1, there, nullptr, o.ctx, -1), "b200asm_exchange_add_peer")) < 0) return rc;
            if (rc != here) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: internal: link index");
            if ((rc = pass(m, o.ctx, b200asm_exchange_add_peer(o.ctx, 0, here, nullptr, p.ctx, in_min), "b200asm_exchange_add_peer")) < 0) return rc;
            if (rc != there) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: internal: link index");
            if ((rc = pass(m, p.ctx, b200asm_exchange_set_map(p.ctx, here, (int64_t)(a1 - a0), p.nnz_owned + (int64_t)a0, a_dst.data(),
                                                               (int64_t)(r1 - r0), rhs_src.data(), rhs_dst.data()),
                           "b200asm_exchange_set_map")) < 0)
                return rc;
            a0 = a1; r0 = r1;
        }
        if (a0 != rem.size() || r0 != rrows.size()) return mfail(m, B200ASM_EINVAL, "b200asm_multi_set_pattern: internal: staged entries without an owner");
    }
    // halo maps of the sharded solver: owner and position there of every halo equation (every own_first is known now)
    for (int k = 0; k < ndev; k++) {
        Part &p = m->parts[k];
        p.halo_owner.resize(p.halo_local.size());
        p.halo_remote.resize(p.halo_local.size());
        for (size_t i = 0; i < p.halo_local.size(); i++) {
            const int64_t eq = p.eqs[p.halo_local[i]];
            const int o = owner_of(eq);
            p.halo_owner[i] = o;
            p.halo_remote[i] = (int32_t)(m->parts[o].own_first + (eq - m->parts[o].row0));
        }
    }
    m->have_pattern = true;
    return 0;
}

// Conjugate gradients on the row-sharded matrix of the last b200asm_multi_assemble (csrc/cg_device.cuh): the reference's CG
// (Solvers/LinearSolvers/cg.h:44-120), the product of TPZSYsmpMatrix::MultAdd (Matrix/pzsysmp.cpp:190-232) split by row blocks,
// halos of p and q over NVLink.  Same arguments as b200asm_cg_solve; f_host / x_host are GLOBAL vectors.
extern "C" int b200asm_multi_cg_solve(b200asm_multi *m, int precond, int64_t max_iter, double tol, int from_current, const double *f_host,
                                      double *x_host, int64_t *iters_out, double *resid_out) {
    if (!m) return B200ASM_EINVAL;
    if (!m->have_pattern) return mfail(m, B200ASM_ESTATE, "b200asm_multi_cg_solve: assemble a matrix first");
    std::vector<b200asm_cg_shard> sh(m->parts.size());
    for (size_t k = 0; k < m->parts.size(); k++) {
        Part &p = m->parts[k];
        sh[k] = {p.ctx, p.own_first, p.row1 - p.row0, p.row0, (int64_t)p.halo_local.size(), p.halo_local.data(), p.halo_owner.data(), p.halo_remote.data()};
    }
    std::string err;
    const int rc = b200asm_cg_sharded((int)sh.size(), sh.data(), precond, max_iter, tol, from_current, f_host, x_host, iters_out, resid_out, err);
    if (rc < 0) return mfail(m, rc, err);
    return 0;
}

extern "C" int b200asm_multi_partition(const b200asm_multi *m, int64_t *row_begin, int64_t *elements, int64_t *staged_entries) {
    if (!m || !m->have_pattern) return B200ASM_ESTATE;
    for (size_t k = 0; k < m->parts.size(); k++) {
        if (row_begin) row_begin[k] = m->parts[k].row0;
        if (elements) elements[k] = m->parts[k].nelements;
        if (staged_entries) staged_entries[k] = m->parts[k].nnz_local - m->parts[k].nnz_owned;
    }
    if (row_begin) row_begin[m->parts.size()] = m->neq;
    return 0;
}

namespace {
// run f(part index) on one host thread per GPU (the reference's strategies also spawn their workers inside Assemble and join
// before returning, pzstrmatrixor.cpp:494-502); every GPU must be driven concurrently: a context waits for its peers' pushes
int for_each_part(b200asm_multi *m, int (*f)(b200asm_multi *, int, void *), void *arg) {
    const int n = (int)m->parts.size();
    std::vector<int> rc(n, 0);
    std::vector<std::thread> th;
    for (int k = 1; k < n; k++) th.emplace_back([&, k]() { rc[k] = f(m, k, arg); });
    rc[0] = f(m, 0, arg);
    for (std::thread &t : th) t.join();
    for (int k = 0; k < n; k++)
        if (rc[k] < 0) return mfail(m, rc[k], "device " + std::to_string(m->parts[k].device) + ": " + b200asm_last_error(m->parts[k].ctx));
    return 0;
}
struct AsmArgs { double *a, *rhs; };
}  // namespace

extern "C" int b200asm_multi_assemble(b200asm_multi *m, double *a_host, double *rhs_host) {
    if (!m) return B200ASM_EINVAL;
    if (!m->have_pattern) return mfail(m, B200ASM_ESTATE, "b200asm_multi_assemble: call b200asm_multi_set_pattern after the last add_group");
    AsmArgs args{a_host, rhs_host};
    return for_each_part(m, [](b200asm_multi *mm, int k, void *arg) {
        const AsmArgs &x = *(const AsmArgs *)arg;
        Part &p = mm->parts[k];
        // every GPU copies its own row block into the caller's global arrays
        return b200asm_assemble(p.ctx, x.a ? x.a + mm->ia[p.row0] : nullptr, x.rhs ? x.rhs + p.row0 : nullptr);
    }, &args);
}

extern "C" int b200asm_multi_assemble_rhs(b200asm_multi *m, double *rhs_host) {
    if (!m) return B200ASM_EINVAL;
    if (!m->have_pattern) return mfail(m, B200ASM_ESTATE, "b200asm_multi_assemble_rhs: call b200asm_multi_set_pattern first (the row partition comes from it)");
    AsmArgs args{nullptr, rhs_host};
    return for_each_part(m, [](b200asm_multi *mm, int k, void *arg) {
        const AsmArgs &x = *(const AsmArgs *)arg;
        Part &p = mm->parts[k];
        int rc = b200asm_assemble_rhs(p.ctx, x.rhs ? x.rhs + p.row0 : nullptr);
        if (rc == 0) rc = b200asm_synchronize(p.ctx);
        return rc;
    }, &args);
}

extern "C" int b200asm_multi_assemble_async(b200asm_multi *m) {
    if (!m) return B200ASM_EINVAL;
    if (!m->have_pattern) return mfail(m, B200ASM_ESTATE, "b200asm_multi_assemble_async: no pattern");
    for (Part &p : m->parts) {  // launches only: one thread may enqueue every GPU's work
        const int rc = pass(m, p.ctx, b200asm_assemble_async(p.ctx), "b200asm_assemble_async");
        if (rc < 0) return rc;
    }
    return 0;
}

extern "C" int b200asm_multi_synchronize(b200asm_multi *m) {
    if (!m) return B200ASM_EINVAL;
    for (Part &p : m->parts) {
        const int rc = pass(m, p.ctx, b200asm_synchronize(p.ctx), "b200asm_synchronize");
        if (rc < 0) return rc;
    }
    return 0;
}

extern "C" int b200asm_multi_counters(const b200asm_multi *m, int64_t *kernel_launches, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!m) return B200ASM_EINVAL;
    int64_t k = 0, h = 0, d = 0;
    for (const Part &p : m->parts) {
        int64_t a = 0, b = 0, c = 0;
        b200asm_counters(p.ctx, &a, &b, &c);
        k += a; h += b; d += c;
    }
    if (kernel_launches) *kernel_launches = k;
    if (h2d_bytes) *h2d_bytes = h;
    if (d2h_bytes) *d2h_bytes = d;
    return 0;
}
