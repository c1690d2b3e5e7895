// Host side of the TB level kernels (level_tb.cuh): table construction and launches.
// Included by snfft.cu inside namespace snfft, after DevLevel / snfft_plan / upload().

// (MB, S, MAXIN = S!) combinations with compiled kernels.  The product uses the first four
// (levels 7..10); the others exist in the host-emulation build so that the CPU test-suite
// can drive every T-phase width through small groups.
#define SNFFT_TB_COMBOS_PRODUCT(X) X(6, 1, 1) X(7, 1, 1) X(7, 2, 2) X(7, 3, 6)
#ifdef SNFFT_HOST_EMU
#define SNFFT_TB_COMBOS(X) SNFFT_TB_COMBOS_PRODUCT(X) X(5, 1, 1) X(5, 2, 2) X(4, 3, 6) X(6, 2, 2) X(5, 3, 6) X(4, 2, 2) X(3, 3, 6)
#else
#define SNFFT_TB_COMBOS(X) SNFFT_TB_COMBOS_PRODUCT(X)
#endif

constexpr int TB_NBUF = 2;          // chains per pass
constexpr int TB_NBUF_ALL = 8;      // inverse steps with s = 1 (k <= 8): every chain in one pass, one CTA per tile
// chains per pass of the inverse kernel at level k: with one fused top sweep the accumulators of all the
// chains fit in registers (k <= 8 vectors), so one CTA handles a whole tile and every parent tile is read once
static int tb_inv_nbuf(int k, int s) {
    const char* e = std::getenv("SNFFT_TB_INV_ALL");       // tuning switch: 0 = always TB_NBUF chains per pass
    if (e && e[0] == '0') return TB_NBUF;
    return (s == 1 && k <= TB_NBUF_ALL) ? TB_NBUF_ALL : TB_NBUF;
}
// CTA shapes: fp32 runs 512 threads (one CTA per SM, <= 128 registers) when the sub-blocks of a tile need
// them and 256 threads otherwise (two CTAs per SM); fp64 needs twice the registers and always runs 256.
constexpr size_t TB_SMEM_CAP_1 = 200 * 1024, TB_SMEM_CAP_2 = 110 * 1024;

// tuning switch (read when a plan is created): SNFFT_TB_NT512=0 restricts fp32 to 256-thread CTAs
static bool tb_nt512() {
    const char* e = std::getenv("SNFFT_TB_NT512");
    return !(e && e[0] == '0');
}
template <typename T>
static int tb_pick_nt(long long b_threads) {
    if (sizeof(T) == 8 || !tb_nt512()) return b_threads <= 256 ? 256 : 0;
    return b_threads <= 256 ? 256 : (b_threads <= 512 ? 512 : 0);
}
template <typename T>
static size_t tb_smem_cap(int nt) {
    return (sizeof(T) == 4 && nt == 256) ? TB_SMEM_CAP_2 : TB_SMEM_CAP_1;
}
static int tb_pad4(int x) { return (x + 3) & ~3; }
// tuning switch: SNFFT_TB_TCMAX = log2 of the widest column-instance tile the TB kernels may use (default 7)
static int tb_tcl_max() {
    const char* e = std::getenv("SNFFT_TB_TCMAX");
    const int v = e ? std::atoi(e) : 7;
    return v < 2 ? 2 : (v > 7 ? 7 : v);
}

static bool tb_combo_compiled(int mb, int s) {
#define SNFFT_TB_HAS(MB, S, MAXIN) \
    if (mb == MB && s == S) return true;
    SNFFT_TB_COMBOS(SNFFT_TB_HAS)
#undef SNFFT_TB_HAS
    return false;
}

// number of fused top sweeps of level k; 0 = this level runs the sweep-per-pass kernels
static int tb_pick_s(int k) {
    const int mode = g_tb_mode.load();
    if (mode == 0) return 0;
    if (mode > 0) return (mode <= TB_MAXS && tb_combo_compiled(k - mode, mode)) ? mode : 0;
    switch (k) {
        case 7: return 1;
        case 8: return 1;
        case 9: return 2;
        case 10: return 3;
        default: return 0;
    }
}

// Directions that use the TB kernels at level k.  Forced modes use them everywhere; the automatic
// choice follows the measurements in profiles/ (r01j): for fp32 they win the forward steps k = 7, 9 and the
// inverse step k = 7 (fp64: forward k = 9); the sweep kernels run the rest.  SNFFT_TB_FWD / SNFFT_TB_INV (lists of levels, e.g.
// "7,8,9") override the automatic choice when a plan is created (tuning only).
static bool tb_level_listed(const char* env, int k, bool dflt) {
    const char* e = std::getenv(env);
    if (!e) return dflt;
    for (const char* q = e; *q;) {
        char* end;
        const long v = std::strtol(q, &end, 10);
        if (end == q) break;
        if (v == k) return true;
        q = (*end == ',') ? end + 1 : end;
    }
    return false;
}
static void tb_pick_dirs(int k, int dtype, bool* fwd, bool* inv) {
    if (g_tb_mode.load() > 0) {
        *fwd = *inv = true;
        return;
    }
    *fwd = tb_level_listed("SNFFT_TB_FWD", k, dtype == SNFFT_F32 ? (k == 7 || k == 9) : k == 9);
    *inv = tb_level_listed("SNFFT_TB_INV", k, dtype == SNFFT_F32 && k == 7);
}

struct TbPath {
    int off, bottom;
};

// all chains of `depth` removals below shape (level, id): row offset of the sub-block and the shape reached
static void tb_paths(const HostPlan& hp, int level, int id, int depth, int off, std::vector<TbPath>& out) {
    if (depth == 0) {
        out.push_back({off, id});
        return;
    }
    const HostShape& S = hp.levels[level].shapes[id];
    for (size_t c = 0; c < S.child.size(); ++c) tb_paths(hp, level - 1, S.child[c], depth - 1, off + S.boff[c], out);
}

template <typename T>
static int build_tb_level(snfft_plan* p, int k, int s) {
    const HostPlan& hp = p->host;
    const HostLevel& L = hp.levels[k];
    const HostLevel& P = hp.levels[k - 1];
    DevLevel& D = p->lev[k];
    const int mb = k - s;
    const int maxin = s == 1 ? 1 : (s == 2 ? 2 : 6);     // coefficient rows are padded to s! entries
    constexpr int VT = 16 / (int)sizeof(T);
    int vtl = 0;
    while ((1 << vtl) < VT) ++vtl;
    const int ns = (int)L.shapes.size();

    std::vector<uint32_t> tt, bt;
    std::vector<T> ct;
    std::vector<std::vector<TbTask>> task_of(ns);        // [lambda][b]
    std::vector<std::vector<int>> tcl_of(ns), nt_of(ns);

    for (int si = 0; si < ns; ++si) {
        const HostShape& S = L.shapes[si];
        std::vector<TbPath> subs, outs;
        tb_paths(hp, k, si, s, 0, subs);                 // sub-blocks: bottoms are partitions of mb
        tb_paths(hp, k, si, s + 1, 0, outs);             // T outputs: bottoms zeta are partitions of mb - 1
        const int nsb = (int)subs.size();
        // threads of one warp run the same generated code when their sub-blocks have the same type;
        // inside a type alternate the parity of the base row (two sub-blocks of a warp then hit different banks)
        // (largest first: the flat B' loop of the inverse kernel then deals balanced rounds)
        {
            const HostLevel& MBL = hp.levels[mb];
            std::stable_sort(subs.begin(), subs.end(), [&](const TbPath& x, const TbPath& y) {
                const int dx = MBL.shapes[x.bottom].dim, dy = MBL.shapes[y.bottom].dim;
                return dx != dy ? dx > dy : x.bottom < y.bottom;
            });
        }
        for (size_t lo = 0; lo < subs.size();) {
            size_t hi = lo;
            while (hi < subs.size() && subs[hi].bottom == subs[lo].bottom) ++hi;
            std::vector<TbPath> ev, od, mix;
            for (size_t q = lo; q < hi; ++q) (subs[q].off & 1 ? od : ev).push_back(subs[q]);
            size_t a = 0, c = 0;
            while (a < ev.size() || c < od.size()) {
                if (a < ev.size()) mix.push_back(ev[a++]);
                if (c < od.size()) mix.push_back(od[c++]);
            }
            std::copy(mix.begin(), mix.end(), subs.begin() + lo);
            lo = hi;
        }
        task_of[si].resize(S.child.size());
        tcl_of[si].resize(S.child.size());
        nt_of[si].resize(S.child.size());
        for (size_t b = 0; b < S.child.size(); ++b) {
            const HostShape& M = P.shapes[S.child[b]];
            std::vector<TbPath> ins;
            tb_paths(hp, k - 1, S.child[b], s, 0, ins);
            std::vector<int> zetas;                      // orbit order: first appearance among the rows of mu
            for (const TbPath& q : ins)
                if (std::find(zetas.begin(), zetas.end(), q.bottom) == zetas.end()) zetas.push_back(q.bottom);
            const int norb = (int)zetas.size();
            tt.resize(tb_pad4((int)tt.size()), 0u);          // 16-byte aligned table starts (cp.async staging)
            ct.resize(tb_pad4((int)ct.size()), (T)0);
            bt.resize(tb_pad4((int)bt.size()), 0u);
            TbTask tk;
            tk.pad = 0;
            tk.lam_off = S.coef_off;
            tk.mu_off = M.coef_off;
            tk.d_lam = S.dim;
            tk.d_mu = M.dim;
            tk.boff = S.boff[b];
            tk.nsb = nsb;
            tk.tt_off = (int)tt.size();
            tk.ct_off = (int)ct.size();
            tk.bt_off = (int)bt.size();
            tk.scale = (double)S.dim / ((double)k * (double)M.dim);

            const size_t t0 = tt.size();
            tt.resize(t0 + TB_TT_HEAD + (size_t)norb * TB_ORB_WORDS + (size_t)(s + 1) * norb * 2, 0u);
            std::vector<std::vector<char>> support(s + 1, std::vector<char>(S.dim, 0));
            int start = 0;
            for (int o = 0; o < norb; ++o) {
                const int len = hp.levels[mb - 1].shapes[zetas[o]].dim;
                std::vector<int> in_rows, out_rows;
                for (const TbPath& q : ins)
                    if (q.bottom == zetas[o]) in_rows.push_back(q.off);
                for (const TbPath& q : outs)
                    if (q.bottom == zetas[o]) out_rows.push_back(q.off);
                const int nin = (int)in_rows.size(), nout_all = (int)out_rows.size();
                if (nin > maxin) return set_error(SNFFT_EINVAL, "TB tables: more inputs per orbit than s!");
                uint32_t* orb = &tt[t0 + TB_TT_HEAD + (size_t)o * TB_ORB_WORDS];
                orb[0] = (uint32_t)start;
                orb[1] = (uint32_t)len;
                orb[2] = (uint32_t)nin;
                for (int q = 0; q < nin; ++q) orb[3 + q] = (uint32_t)in_rows[q];
                start += len;
                // columns of the class matrices: apply the top sweeps to the unit vectors of the inputs
                // (rows at t = 0 of the orbit; rho(s_j), j >= k-1-s, never leaves the orbit)
                std::vector<int> pos(S.dim, -1);
                for (int e = 0; e < nout_all; ++e) pos[out_rows[e]] = e;
                std::vector<std::vector<double>> v(nin, std::vector<double>(nout_all, 0.0));
                std::vector<char> sup(nout_all, 0);
                for (int q = 0; q < nin; ++q) {
                    const int e = pos[S.boff[b] + in_rows[q]];
                    if (e < 0) return set_error(SNFFT_EINVAL, "TB tables: input row is not an orbit row");
                    v[q][e] = 1.0;
                    sup[e] = 1;
                }
                for (int cls = 0; cls <= s; ++cls) {
                    if (cls > 0) {
                        const int j = k - 1 - cls;       // sweeps k-2, k-3, ... in the order they are applied
                        std::vector<char> nsup(sup);
                        for (int e = 0; e < nout_all; ++e) {
                            const int pr = S.partner[j][out_rows[e]];
                            if (pr != out_rows[e]) {
                                if (pos[pr] < 0) return set_error(SNFFT_EINVAL, "TB tables: sweep leaves the orbit");
                                if (sup[pos[pr]]) nsup[e] = 1;
                            }
                        }
                        for (int q = 0; q < nin; ++q) {
                            std::vector<double> nv(nout_all, 0.0);
                            for (int e = 0; e < nout_all; ++e) {
                                const int row = out_rows[e];
                                const int pr = S.partner[j][row];
                                double acc = S.diag[j][row] * v[q][e];
                                if (pr != row) acc += S.off[j][row] * v[q][pos[pr]];
                                nv[e] = acc;
                            }
                            v[q].swap(nv);
                        }
                        sup.swap(nsup);
                    }
                    uint32_t* dir = &tt[t0 + TB_TT_HEAD + (size_t)norb * TB_ORB_WORDS + ((size_t)cls * norb + o) * 2];
                    dir[0] = (uint32_t)(tt.size() - t0);
                    dir[1] = (uint32_t)(ct.size() - (size_t)tk.ct_off);
                    std::vector<uint32_t> ol(1, 0u);
                    for (int e = 0; e < nout_all; ++e) {
                        if (!sup[e]) continue;
                        ol.push_back((uint32_t)out_rows[e]);
                        for (int q = 0; q < maxin; ++q) ct.push_back(q < nin ? (T)v[q][e] : (T)0);
                        for (int t = 0; t < len; ++t) support[cls][out_rows[e] + t] = 1;
                    }
                    ol[0] = (uint32_t)(ol.size() - 1);
                    // `dir` may dangle after the insert: it was written before
                    tt.insert(tt.end(), ol.begin(), ol.end());
                }
            }
            tt[t0 + 0] = (uint32_t)norb;
            tt[t0 + 1] = (uint32_t)start;
            tt[t0 + 2] = (uint32_t)s;
            tt[t0 + 3] = (uint32_t)(tt.size() - t0);
            if (norb > 255) return set_error(SNFFT_ENOTSUP, "TB tables: too many orbits");
            {
                std::vector<uint32_t> r2o((start + 3) / 4, 0u);
                for (int o = 0; o < norb; ++o) {
                    const uint32_t* orb = &tt[t0 + TB_TT_HEAD + (size_t)o * TB_ORB_WORDS];
                    for (uint32_t r = orb[0]; r < orb[0] + orb[1]; ++r) r2o[r / 4] |= (uint32_t)o << (8 * (r % 4));
                }
                tt.insert(tt.end(), r2o.begin(), r2o.end());
            }
            tk.R = start;
            tt.resize(tb_pad4((int)tt.size()), 0u);
            ct.resize(tb_pad4((int)ct.size()), (T)0);
            tk.tt_words = (int)(tt.size() - t0);
            tk.ct_len = (int)(ct.size() - (size_t)tk.ct_off);
            // sub-block table with the row masks of every class
            for (const TbPath& sbp : subs) {
                const int d = hp.levels[mb].shapes[sbp.bottom].dim;
                if (sbp.off > 0xffff || sbp.bottom > 0xffff || d > 64)
                    return set_error(SNFFT_ENOTSUP, "TB tables: sub-block does not fit the table fields");
                bt.push_back((uint32_t)sbp.off | ((uint32_t)sbp.bottom << 16));
                for (int cls = 0; cls <= TB_MAXS; ++cls) {
                    uint64_t mask = 0;
                    if (cls <= s)
                        for (int r = 0; r < d; ++r)
                            if (support[cls][sbp.off + r]) mask |= (uint64_t)1 << r;
                    bt.push_back((uint32_t)(mask & 0xffffffffu));
                    bt.push_back((uint32_t)(mask >> 32));
                }
            }
            bt.resize(tb_pad4((int)bt.size()), 0u);
            tk.bt_words = (int)(bt.size() - (size_t)tk.bt_off);
            // tile width: the largest whose B role fits a CTA and whose buffers fit shared memory
            int tcl = tb_tcl_max(), nt = 0;
            for (; tcl >= vtl; --tcl) {
                nt = tb_pick_nt<T>((long long)nsb << tcl);
                if (nt && tb_smem_bytes<T>(false, TB_NBUF, S.dim << tcl, M.dim << tcl, tk.ct_len, tk.tt_words, tk.bt_words) <=
                              tb_smem_cap<T>(nt))
                    break;
            }
            if (tcl < vtl) return SNFFT_ENOTSUP;          // caller falls back to the sweep kernels
            task_of[si][b] = tk;
            tcl_of[si][b] = tcl;
            nt_of[si][b] = nt;
        }
    }

    // forward ranges: tasks grouped by tile width, big columns first
    std::vector<TbTask> tasks;
    for (int key = 15; key >= 0; --key) {
        const int tcl = key >> 1, nt = (key & 1) ? 512 : 256;
        std::vector<TbTask> grp;
        for (int si = 0; si < ns; ++si)
            for (size_t b = 0; b < task_of[si].size(); ++b)
                if (tcl_of[si][b] == tcl && nt_of[si][b] == nt) grp.push_back(task_of[si][b]);
        if (grp.empty()) continue;
        std::stable_sort(grp.begin(), grp.end(), [](const TbTask& x, const TbTask& y) { return x.d_lam > y.d_lam; });
        TbRange r{};
        r.tcl = tcl;
        r.nt = nt;
        r.begin = (int)tasks.size();
        r.count = (int)grp.size();
        for (const TbTask& t : grp) {
            r.elems += (long long)t.d_lam * t.d_mu;
            r.max_dmu = std::max(r.max_dmu, t.d_mu);
            r.max_dlam = std::max(r.max_dlam, t.d_lam);
            r.max_tt = std::max(r.max_tt, t.tt_words);
            r.max_ct = std::max(r.max_ct, t.ct_len);
            r.max_bt = std::max(r.max_bt, t.bt_words);
            r.max_nsb = std::max(r.max_nsb, t.nsb);
            tasks.push_back(t);
        }
        D.tb_franges.push_back(r);
    }
    // inverse tasks: one per mu; its parents are appended to the task array
    std::vector<TbInvTask> itasks_all;
    std::vector<int> itcl, int_;
    std::vector<std::array<int, 4>> imax;       // max d_lam, tt, ct, bt over the parents
    for (size_t m = 0; m < P.shapes.size(); ++m) {
        const HostShape& M = P.shapes[m];
        TbInvTask it;
        it.mu_off = M.coef_off;
        it.d_mu = M.dim;
        it.par_off = (int)tasks.size();
        it.npar = (int)M.parent.size();
        it.R = 0;
        int maxd = 0, maxnsb = 0, mtt = 0, mct = 0, mbt = 0;
        for (size_t q = 0; q < M.parent.size(); ++q) {
            const TbTask& t = task_of[M.parent[q]][M.parent_block[q]];
            if (q > 0 && t.R != it.R) return set_error(SNFFT_EINVAL, "TB tables: orbit rows differ between parents");
            it.R = t.R;
            maxd = std::max(maxd, t.d_lam);
            maxnsb = std::max(maxnsb, t.nsb);
            mtt = std::max(mtt, t.tt_words);
            mct = std::max(mct, t.ct_len);
            mbt = std::max(mbt, t.bt_words);
            tasks.push_back(t);
        }
        int tcl = tb_tcl_max(), nt = 0;
        for (; tcl >= vtl; --tcl) {
            nt = tb_pick_nt<T>(std::max((long long)maxnsb << tcl, (long long)it.R << (tcl - vtl)));
            if (nt && tb_smem_bytes<T>(true, tb_inv_nbuf(k, s), maxd << tcl, 0, mct, mtt, mbt) <= tb_smem_cap<T>(nt)) break;
        }
        if (tcl < vtl) return SNFFT_ENOTSUP;
        itasks_all.push_back(it);
        itcl.push_back(tcl);
        int_.push_back(nt);
        imax.push_back({maxd, mtt, mct, mbt});
    }
    std::vector<TbInvTask> itasks;
    for (int key = 15; key >= 0; --key) {
        const int tcl = key >> 1, nt = (key & 1) ? 512 : 256;
        TbRange r{};
        r.tcl = tcl;
        r.nt = nt;
        r.begin = (int)itasks.size();
        for (size_t m = 0; m < itasks_all.size(); ++m) {
            if (itcl[m] != tcl || int_[m] != nt) continue;
            const TbInvTask& it = itasks_all[m];
            r.elems += (long long)k * it.d_mu * it.d_mu;
            r.max_dmu = std::max(r.max_dmu, it.d_mu);
            r.max_dlam = std::max(r.max_dlam, imax[m][0]);
            r.max_tt = std::max(r.max_tt, imax[m][1]);
            r.max_ct = std::max(r.max_ct, imax[m][2]);
            r.max_bt = std::max(r.max_bt, imax[m][3]);
            itasks.push_back(it);
            ++r.count;
        }
        if (r.count) D.tb_iranges.push_back(r);
    }

    int rc;
    if ((rc = upload(p, tasks, &D.tb_tasks))) return rc;
    if ((rc = upload(p, itasks, &D.tb_itasks))) return rc;
    if ((rc = upload(p, tt, &D.tb_tt))) return rc;
    if ((rc = upload(p, ct, &D.tb_ct))) return rc;
    if ((rc = upload(p, bt, &D.tb_bt))) return rc;
    D.tb_s = s;
    D.tb_inv_nbuf = tb_inv_nbuf(k, s);
    return SNFFT_OK;
}

template <typename T>
static TbArgs<T> tb_args(const snfft_plan* p, int k, const void* in, void* out, long long B, const TbRange& r) {
    const DevLevel& D = p->lev[k];
    constexpr int VT = 16 / (int)sizeof(T);
    TbArgs<T> a;
    a.in = (const T*)in;
    a.out = (T*)out;
    a.tasks = (const TbTask*)D.tb_tasks;
    a.itasks = (const TbInvTask*)D.tb_itasks;
    a.tt = (const uint32_t*)D.tb_tt;
    a.ct = (const T*)D.tb_ct;
    a.bt = (const uint32_t*)D.tb_bt;
    a.k = k;
    a.s = D.tb_s;
    a.tc_log2 = r.tcl;
    a.bstride = r.max_dlam << r.tcl;
    a.xstride = r.max_dmu << r.tcl;
    a.ct_smem_len = tb_pad4(r.max_ct);                   // keeps every table 16-byte aligned
    a.tt_smem_words = tb_pad4(r.max_tt);
    a.bt_smem_words = tb_pad4(r.max_bt);
    a.ninst = B * (p->host.factorial[p->n] / p->host.factorial[k]);
    a.vec = (a.ninst % VT == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
    {
        const char* e = std::getenv("SNFFT_TB_SKIP");
        a.skip = e ? std::atoi(e) : 0;
    }
    return a;
}

template <typename T, bool INVERSE>
static int run_tb_level(const snfft_plan* p, int k, const void* in, void* out, long long B, cudaStream_t st) {
    const DevLevel& D = p->lev[k];
    const int s = D.tb_s, mb = k - s;
    for (const TbRange& r : INVERSE ? D.tb_iranges : D.tb_franges) {
        TbArgs<T> a = tb_args<T>(p, k, in, out, B, r);
        if (INVERSE) a.itasks += r.begin;
        else a.tasks += r.begin;
        const int nbuf = INVERSE ? D.tb_inv_nbuf : TB_NBUF;
        const size_t smem = tb_smem_bytes<T>(INVERSE, nbuf, a.bstride, a.xstride, a.ct_smem_len, a.tt_smem_words,
                                             a.bt_smem_words);
        const long long tiles = (((long long)r.max_dmu * a.ninst) + (1ll << r.tcl) - 1) >> r.tcl;
        // forward: persistent CTAs, about SNFFT_TB_WAVES (default 8) CTAs per SM slot in flight over the whole range
        long long gx = tiles * ((k + nbuf - 1) / nbuf);
        if (!INVERSE) {
            const char* e = std::getenv("SNFFT_TB_WAVES");
            const long long waves = e ? std::atoll(e) : 8;
            const long long slots = 148ll * (r.nt == 256 ? 2 : 1) * std::max(1ll, waves);
            gx = waves <= 0 ? tiles : std::max(1ll, std::min(tiles, (slots + r.count - 1) / r.count));
            const char* cap = std::getenv("SNFFT_TB_GX");      // test hook: explicit cap of the tile dimension
            if (cap && std::atoll(cap) > 0) gx = std::min(gx, (long long)std::atoll(cap));
        }
        if (gx > 0x7fffffffll) return set_error(SNFFT_ENOTSUP, "batch too large for one launch");
        dim3 grid((unsigned)gx, (unsigned)r.count, 1), block((unsigned)r.nt, 1, 1);
        ProfScope prof(p, st, INVERSE ? 1 : 0, k, 10 * (r.nt / 256) + r.tcl, r.elems);
        int rc = SNFFT_ENOTSUP;
#define SNFFT_TB_LAUNCH(NT_, MINB_, MB, MAXIN)                                  \
    do {                                                                        \
        if (INVERSE) {                                                          \
            if (MAXIN == 1 && nbuf == TB_NBUF_ALL) {                            \
                auto kfn = inv_tb_kernel<T, NT_, MINB_, MB, 1, TB_NBUF_ALL>;    \
                if ((rc = set_smem(kfn, smem))) return rc;                      \
                SNFFT_LAUNCH(kfn, grid, block, smem, st, a);                    \
            } else {                                                            \
                auto kfn = inv_tb_kernel<T, NT_, MINB_, MB, MAXIN, TB_NBUF>;    \
                if ((rc = set_smem(kfn, smem))) return rc;                      \
                SNFFT_LAUNCH(kfn, grid, block, smem, st, a);                    \
            }                                                                   \
        } else {                                                                \
            auto kfn = fwd_tb_kernel<T, NT_, MINB_, MB, MAXIN, TB_NBUF>;        \
            if ((rc = set_smem(kfn, smem))) return rc;                          \
            SNFFT_LAUNCH(kfn, grid, block, smem, st, a);                        \
        }                                                                       \
        rc = SNFFT_OK;                                                          \
    } while (0)
#define SNFFT_TB_CASE(MB, S, MAXIN)                                             \
    if (mb == MB && s == S) {                                                   \
        if constexpr (sizeof(T) == 8) {                                         \
            SNFFT_TB_LAUNCH(256, 1, MB, MAXIN);                                 \
        } else if (r.nt == 512) {                                               \
            SNFFT_TB_LAUNCH(512, 1, MB, MAXIN);                                 \
        } else {                                                                \
            SNFFT_TB_LAUNCH(256, 2, MB, MAXIN);                                 \
        }                                                                       \
    }
        SNFFT_TB_COMBOS(SNFFT_TB_CASE)
#undef SNFFT_TB_CASE
#undef SNFFT_TB_LAUNCH
        if (rc) return set_error(rc, "no TB kernel for this (MB, s)");
        g_launches.fetch_add(1);
        SNFFT_CUDA(cudaGetLastError());
    }
    return SNFFT_OK;
}
