// Host side of a refactorization (BasisSolver::reset, solver.rs:1286-1303): arena capacities, the compact row copy of the basic
// columns, pinned staging, index sets (from scratch or derived from the previous ones), the product-form refresh of the core
// inverse with its accuracy probe, and refactor_impl itself.  Included by engine.cu only, after the launch helpers.
#pragma once

// Column cache / LU arenas.  First allocation is generous (~1 GB of basis columns): cudaFree/cudaMalloc of the big
// arenas costs tens of milliseconds, so capacity grows by doubling and rarely; the cache content survives growth.
static mlp_status ensure_lu_capacity(mlp_engine* e, int64_t k, bool exact = false) {
  if (k <= e->kcap && e->Bcols) return MLP_OK;
  int64_t cap = std::max<int64_t>(e->kcap, std::min<int64_t>(e->m, std::max<int64_t>(64, std::min<int64_t>(1024, (1ll << 30) / (8 * e->mld)))));
  // sparse storage keeps no column cache: the arenas are kcap^2 (factors, inverse) — start at 4096 columns (0.27 GB) so that the
  // first thousands of pivots meet no growth (each growth is a re-allocation AND a true factorization: 5 - 20 ms on config 4)
  if (e->sparse) cap = std::max<int64_t>(e->kcap, std::min<int64_t>(e->m, 4096));
  while (cap < k) cap *= 2;  // may exceed m: slots of columns that left since the last refactor stay occupied
  if (exact) cap = k;        // clone: same leading dimensions as the source
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  PoolScope pool(e->use_pool ? e->stream : nullptr);
  double* nb = nullptr;
  ST(dev_alloc(&nb, e->sparse ? 1 : (size_t)e->mld * cap));  // sparse storage reads the basic columns from the matrix itself
  if (!e->sparse && e->Bcols && e->kcap > 0) {
    CU(cudaMemcpyAsync(nb, e->Bcols, (size_t)e->mld * e->kcap * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  for (int64_t s = cap - 1; s >= e->kcap; --s) e->h_free_slots.push_back((int32_t)s);
  dev_free(e->Jpos); dev_free(e->Jslot); dev_free(e->Rp); dev_free(e->Bcols); dev_free(e->LUc); dev_free(e->Cinv);
  dev_free(e->lu_aff); dev_free(e->lu_perm); dev_free(e->lu_rcnt);
  ST(dev_alloc(&e->lu_aff, 192)); ST(dev_alloc(&e->lu_perm, cap)); ST(dev_alloc(&e->lu_rcnt, cap));
  if (e->sparse) {
    if (e->corevar_k > 0) {  // un-mark with the old list before it is freed
      LAUNCH(e, k_set_corepos, cdiv(e->corevar_k, 256), 256, 0, e->corepos, e->corevar, (int)e->corevar_k, 1);
      CU(cudaStreamSynchronize(e->stream));
      e->corevar_k = 0;
    }
    dev_free(e->corevar); dev_free(e->cseg_first);
    ST(dev_alloc(&e->corevar, cap)); ST(dev_alloc(&e->cseg_first, cap + 1));
    dev_free(e->rf_map); dev_free(e->rf_W); dev_free(e->rf_T); dev_free(e->rf_Ep);
    ST(dev_alloc(&e->rf_map, 3 * (size_t)cap + 2 * RF_CAP));
    ST(dev_alloc(&e->rf_W, (size_t)RF_CAP * cap)); ST(dev_alloc(&e->rf_T, (size_t)RF_CAP * cap)); ST(dev_alloc(&e->rf_Ep, (size_t)RF_CAP * cap));
    e->inv_valid = false;  // C^-1 does not survive the re-allocation: the next refactorization is a true one
  }
  e->Bcols = nb;
  e->kcap = cap;
  ST(dev_alloc(&e->Jpos, cap)); ST(dev_alloc(&e->Jslot, cap)); ST(dev_alloc(&e->Rp, cap));
  ST(dev_alloc(&e->LUc, (size_t)cap * cap)); ST(dev_alloc(&e->Cinv, (size_t)cap * cap));
  for (int l = 0; l < 2; ++l) {
    Lane& ln = e->lane[l];
    dev_free(ln.xk); dev_free(ln.xk2); dev_free(ln.gt_part_k);
    ST(dev_alloc(&ln.xk, cap)); ST(dev_alloc(&ln.xk2, cap)); ST(dev_alloc(&ln.gt_part_k, (size_t)GT_MAXSPLIT * cap));
  }
  return MLP_OK;
}
static mlp_status ensure_eta_capacity(mlp_engine* e, int64_t K, bool exact = false) {
  if (K <= e->Kcap && e->E) return MLP_OK;
  int64_t cap = std::max<int64_t>(e->Kcap, std::max<int64_t>(96, std::min<int64_t>(2080, (2ll << 30) / (8 * e->mld))));
  while (cap < K) cap *= 2;
  if (exact) cap = K;
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  PoolScope pool(e->use_pool ? e->stream : nullptr);
  dev_free(e->E); dev_free(e->Ginv); dev_free(e->gK); dev_free(e->etaR); dev_free(e->etaPrev); dev_free(e->etaHead);
  e->Kcap = cap;
  ST(dev_alloc(&e->E, (size_t)e->mld * cap)); ST(dev_alloc(&e->Ginv, (size_t)cap * cap)); ST(dev_alloc(&e->gK, cap));
  ST(dev_alloc(&e->etaR, cap)); ST(dev_alloc(&e->etaPrev, cap)); ST(dev_alloc(&e->etaHead, cap));
  for (int l = 0; l < 2; ++l) {
    Lane& ln = e->lane[l];
    dev_free(ln.tK); dev_free(ln.tK2); dev_free(ln.gt_part_K);
    ST(dev_alloc(&ln.tK, cap)); ST(dev_alloc(&ln.tK2, cap)); ST(dev_alloc(&ln.gt_part_K, (size_t)GT_MAXSPLIT * cap));
  }
  return MLP_OK;
}

// Compact row-major copy of the k basic structural columns (core column ids), built on the device from the CSC copy with the
// same segmented counting transpose as the CSC copy itself (sparse_build.cuh; input "rows" = the core columns in core
// order, so every row of the copy lists its entries in ascending core column).  jvar: the core columns' variables, already
// uploaded to e->corevar.
constexpr int DCSR_CHUNKS = 128;
static mlp_status build_core_rows(mlp_engine* e, const std::vector<int32_t>& jvar) {
  const int64_t m = e->m, k = (int64_t)jvar.size();
  PoolScope pool(e->use_pool ? e->stream : nullptr);
  if (!e->dcsr_ptr) {
    ST(dev_alloc(&e->dcsr_ptr, (size_t)e->mld + 1));
    ST(dev_alloc(&e->dcsr_hist, (size_t)DCSR_CHUNKS * e->mld));
    ST(dev_alloc(&e->dcsr_cnt, (size_t)e->mld));
  }
  if (k == 0) {
    CU(cudaMemsetAsync(e->dcsr_ptr, 0, (size_t)(m + 1) * sizeof(int64_t), e->stream));
    return MLP_OK;
  }
  int64_t nz = 0;
  for (int32_t v : jvar) nz += e->h_csc_ptr[(size_t)v + 1] - e->h_csc_ptr[(size_t)v];
  if (nz > e->dcsr_cap) {
    CU(cudaStreamSynchronize(e->lane[1].st));
    dev_free(e->dcsr_idx); dev_free(e->dcsr_val);
    e->dcsr_cap = std::max<int64_t>(4 * nz, 1 << 22);  // 12 bytes per entry: grow rarely
    ST(dev_alloc(&e->dcsr_idx, (size_t)e->dcsr_cap)); ST(dev_alloc(&e->dcsr_val, (size_t)e->dcsr_cap));
  }
  const int ncseg = (int)e->ncseg;                                   // the core's segments (e->cseg_id), in core-column order
  const int spc = (ncseg + DCSR_CHUNKS - 1) / DCSR_CHUNKS;          // segments per chunk
  const int chunks = (ncseg + spc - 1) / spc;
  CU(cudaMemsetAsync(e->dcsr_hist, 0, (size_t)chunks * m * sizeof(int32_t), e->stream));
  LAUNCH(e, k_d_hist, cdiv((int64_t)ncseg * 32, 256), 256, 0, (const int4*)e->seg_desc, e->cseg_id, ncseg, e->csc_idx, m, spc, e->dcsr_hist);
  // per constraint row: scan over the chunks + row counts (k_t_colscan's segment output is not needed: lane scratch)
  LAUNCH(e, k_t_colscan, cdiv(m, 256), 256, 0, e->dcsr_hist, m, chunks, e->dcsr_cnt, (int64_t*)e->lane[0].wm, CSC_SEG);
  LAUNCH(e, k_scan_excl, 1, 1024, 0, e->dcsr_cnt, m, e->dcsr_ptr);
  LAUNCH(e, k_d_fill, chunks, 256, 0, (const int4*)e->seg_desc, e->cseg_id, ncseg, e->csc_idx, e->csc_val, m, spc, e->corepos, e->dcsr_hist,
         e->dcsr_ptr, e->dcsr_idx, e->dcsr_val);
  return MLP_OK;
}

static void refac_stage(mlp_engine* e, const char* name) {
  if (!e->refac_trace) return;
  if (e->refac_trace == 1) for (int l = 0; l < 2; ++l) cudaStreamSynchronize(e->lane[l].st);  // 2: host-side times only, no extra syncs
  const auto now = std::chrono::steady_clock::now();
  if (name) {
    const double ms = std::chrono::duration<double, std::milli>(now - e->refac_t).count();
    if (e->refac_trace == 2 && ms > 0.5)
      fprintf(stderr, "[refactor event] #%lld since-lu %lld k %lld K %lld: %.3f ms in '%s'\n", (long long)e->cnt.refactors,
              (long long)e->pivots_since_lu, (long long)e->k, (long long)e->K, ms, name);
    bool found = false;
    for (auto& st : e->refac_stage) if (st.first == name) { st.second += ms; found = true; break; }
    if (!found) e->refac_stage.emplace_back(name, ms);
  }
  e->refac_t = std::chrono::steady_clock::now();
}
static void refac_report(mlp_engine* e) {
  if (!e->refac_trace || e->cnt.refactors == 0) return;
  double tot = 0.0;
  for (auto& st : e->refac_stage) tot += st.second;
  fprintf(stderr, "[refactor trace] %lld refactorizations (%lld of them product-form refreshes), mean k %.0f, %.3f ms each\n",
          (long long)e->cnt.refactors, (long long)e->cnt.refreshes, e->refac_k_sum / e->cnt.refactors, tot / e->cnt.refactors);
  fprintf(stderr, "[refactor trace] host blocked in %lld per-pivot device waits: %.1f ms in total (%.1f us each) over %lld basis changes, %lld launches\n",
          (long long)e->waits, e->wait_ms, e->waits ? 1e3 * e->wait_ms / e->waits : 0.0, (long long)e->pivot_seq, (long long)e->cnt.kernel_launches);
  fprintf(stderr, "[refactor trace] accuracy probe (normwise backward error of sampled columns of C^-1): worst accepted refresh %.3g, "
                  "worst after a true factorization %.3g, tolerance %.3g, rejected refreshes %lld\n", e->rf_worst, e->rf_worst_true, e->rf_tol,
          (long long)e->cnt.refresh_rejects);
  for (auto& st : e->refac_stage) fprintf(stderr, "[refactor trace]   %-28s %9.3f ms each  %5.1f %%\n", st.first, st.second / e->cnt.refactors, 100.0 * st.second / tot);
}

// Product-form refresh (refresh_inverse.cuh): C_new^-1 from C_old^-1 and the eta file, written into the LUc buffer, which
// then becomes Cinv.  jpos / R: the NEW core's positions and rows.  Runs before anything of the old factor state (index maps,
// compact core rows, eta file) is touched; both lanes are drained.
// Pinned staging: every index array of a refactorization goes through ONE pinned buffer (a cudaMemcpyAsync from pageable memory
// first waits for the stream and then copies synchronously: eight of them serialised the host with the device).
static mlp_status stage_begin(mlp_engine* e, size_t ints_needed) {
  e->stg_cur ^= 1;
  const int c = e->stg_cur;
  if (!e->stg_ev[c]) CU(cudaEventCreateWithFlags(&e->stg_ev[c], cudaEventDisableTiming));
  else CU(cudaEventSynchronize(e->stg_ev[c]));  // the copies that last used this buffer (two refactorizations ago) are long done
  if (ints_needed > e->stg_cap[c]) {
    if (e->stg_h[c]) cudaFreeHost(e->stg_h[c]);
    e->stg_h[c] = nullptr;
    e->stg_cap[c] = std::max<size_t>(2 * ints_needed, (size_t)1 << 16);
    CU(cudaHostAlloc((void**)&e->stg_h[c], e->stg_cap[c] * sizeof(int32_t), cudaHostAllocDefault));
  }
  e->stg_off = 0;
  return MLP_OK;
}
static mlp_status stage_put(mlp_engine* e, void* dst_dev, const int32_t* src, size_t n) {
  if (n == 0) return MLP_OK;
  const int c = e->stg_cur;
  if (e->stg_off + n > e->stg_cap[c]) { set_err("refactor: staging buffer too small"); return MLP_INVALID; }
  int32_t* h = e->stg_h[c] + e->stg_off;
  std::memcpy(h, src, n * sizeof(int32_t));
  e->stg_off += n;
  CU(cudaMemcpyAsync(dst_dev, h, n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  e->cnt.h2d_bytes += (int64_t)(n * sizeof(int32_t));
  return MLP_OK;
}
static mlp_status stage_end(mlp_engine* e) {
  CU(cudaEventRecord(e->stg_ev[e->stg_cur], e->stream));
  return MLP_OK;
}

// dst[idx[i]] = val[i]
__global__ void k_patch_i32(int32_t* __restrict__ dst, const int32_t* __restrict__ idx, const int32_t* __restrict__ val, int n) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = val[i];
}
// rowcore[Rp[c]] = c
__global__ void k_set_rowcore(int32_t* __restrict__ rowcore, const int32_t* __restrict__ Rp, int k) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < k) rowcore[Rp[c]] = c;
}

// The index sets of the new basis derived from those of the factorized one and the list of basis changes since
// (sparse storage, every change recorded): O(k + K log k) instead of three passes over all m positions / rows.
//   jpos   core columns' positions in order_simple's order (ordering.rs:4-21: ascending entry count, ascending position within a count)
//   R      core rows, ascending
//   prow / pval   rows whose rowcover entry changes and the new values (device patch)
//   gone   rows that left the core (rowcore <- -1)
// e->h_rowcover_f is updated in place.
static mlp_status incremental_sets(mlp_engine* e, std::vector<int32_t>& jpos, std::vector<int32_t>& R, std::vector<int32_t>& prow,
                                   std::vector<int32_t>& pval, std::vector<int32_t>& gone) {
  const int64_t ng = e->ng;
  std::vector<int32_t> P;       // distinct changed positions
  std::vector<int64_t> oldv;    // variable the factorized basis held there
  for (size_t j = 0; j < e->h_eta_pos.size(); ++j) {
    const int32_t p = e->h_eta_pos[j];
    if (std::find(P.begin(), P.end(), p) == P.end()) { P.push_back(p); oldv.push_back(e->h_eta_leave[j]); }
  }
  std::vector<int32_t>& rc = e->h_rowcover_f;
  std::vector<int32_t> T;  // slack rows involved
  e->h_rc_old_rows.clear();
  e->h_rc_old_vals.clear();
  auto touch = [&](int32_t i) {
    if (std::find(T.begin(), T.end(), i) != T.end()) return;
    T.push_back(i);
    e->h_rc_old_rows.push_back(i);
    e->h_rc_old_vals.push_back(rc[(size_t)i]);  // the factorized basis' value: the refresh needs it
  };
  for (size_t q = 0; q < P.size(); ++q) if (oldv[q] >= ng) touch((int32_t)(oldv[q] - ng));
  for (size_t q = 0; q < P.size(); ++q) { const int64_t nv = e->h_bvar[(size_t)P[q]]; if (nv >= ng) touch((int32_t)(nv - ng)); }
  for (size_t q = 0; q < P.size(); ++q) if (oldv[q] >= ng) rc[(size_t)(oldv[q] - ng)] = -1;
  for (size_t q = 0; q < P.size(); ++q) { const int64_t nv = e->h_bvar[(size_t)P[q]]; if (nv >= ng) rc[(size_t)(nv - ng)] = P[q]; }
  auto key_less = [&](int32_t pa, int32_t pb) {  // order_simple's key of the column at a position
    const int64_t va = e->h_bvar[(size_t)pa], vb = e->h_bvar[(size_t)pb];
    const int64_t ca = e->h_csc_ptr[(size_t)va + 1] - e->h_csc_ptr[(size_t)va], cb = e->h_csc_ptr[(size_t)vb + 1] - e->h_csc_ptr[(size_t)vb];
    return ca != cb ? ca < cb : pa < pb;
  };
  jpos.clear();
  jpos.reserve(e->h_Jpos_f.size() + P.size());
  std::vector<int32_t> Ps(P);
  std::sort(Ps.begin(), Ps.end());
  for (int32_t p : e->h_Jpos_f)
    if (!std::binary_search(Ps.begin(), Ps.end(), p)) jpos.push_back(p);  // unchanged columns keep their relative order
  for (size_t q = 0; q < P.size(); ++q) {
    if (e->h_bvar[(size_t)P[q]] >= ng) continue;
    jpos.insert(std::lower_bound(jpos.begin(), jpos.end(), P[q], key_less), P[q]);
  }
  R = e->h_R_sorted;
  prow.clear(); pval.clear(); gone.clear();
  for (int32_t i : T) {
    prow.push_back(i);
    pval.push_back(rc[(size_t)i]);
    auto it = std::lower_bound(R.begin(), R.end(), i);
    const bool in_old = it != R.end() && *it == i;
    const bool in_new = rc[(size_t)i] < 0;
    if (in_old && !in_new) { R.erase(it); gone.push_back(i); }
    else if (!in_old && in_new) R.insert(it, i);
  }
  return MLP_OK;
}

// k_rf_probe on the current C^-1 against the compact rows of the current basic columns; one read-back.  *err: the largest
// normwise backward error over the sampled columns, *core_entries: entries of the core.
static mlp_status probe_inverse(mlp_engine* e, int64_t k, double* err, int64_t* core_entries) {
  unsigned long long* out = e->d_nnzcnt;
  CU(cudaMemsetAsync(out, 0, (1 + 2 * RF_PROBE) * sizeof(unsigned long long), e->stream));
  const int ncol = (int)std::min<int64_t>(RF_PROBE, k);
  LAUNCH(e, k_rf_probe, cdiv(k, 256), 256, 0, e->dcsr_ptr, e->dcsr_idx, e->dcsr_val, e->Rp, (int)k, e->Cinv, e->kcap,
         (int)((e->cnt.refactors * 2654435761ull) % (unsigned long long)k), (int)std::max<int64_t>(1, k / RF_PROBE), ncol, out);
  unsigned long long h[1 + 2 * RF_PROBE];
  ST(d2h(e, h, out, sizeof(h)));
  *core_entries = (int64_t)h[0];
  double worst = 0.0;
  for (int q = 0; q < ncol; ++q) {
    double num, den;
    std::memcpy(&num, &h[1 + q], 8);
    std::memcpy(&den, &h[1 + RF_PROBE + q], 8);
    const double r = den > 0.0 ? num / den : num;
    if (!(r <= worst)) worst = r;
  }
  *err = worst;
  return MLP_OK;
}
static bool can_refresh(const mlp_engine* e) {
  // the rank-K product costs 2 k^2 K flops: worth it while the eta file is short next to the core (a factorization is ~2 k^3)
  return e->sparse && e->inv_valid && e->lu_every > 0 && e->k > 0 && e->K >= 1 && e->K <= std::min<int64_t>(e->Kcap, RF_CAP) &&
         e->K <= std::max<int64_t>(RF_MAXK, e->k / 2) &&
         (int64_t)e->h_Jpos_f.size() == e->k && (int64_t)e->h_eta_pos.size() == e->K && (int64_t)e->h_pos_core.size() == e->m &&
         e->rf_map != nullptr;
}
static mlp_status refresh_inverse(mlp_engine* e, const std::vector<int32_t>& jpos, const std::vector<int32_t>& R) {
  const int k_old = (int)e->k, K = (int)e->K, k_new = (int)jpos.size();
  const int64_t ld = e->kcap;
  std::vector<int32_t> qpos, wrow;  // old slack positions whose row of B_old^-1 is needed, and the rows of those slacks
  auto q_of = [&](int32_t p) -> int {
    for (size_t q = 0; q < qpos.size(); ++q) if (qpos[q] == p) return (int)q;
    return -1;
  };
  std::vector<int32_t> map((size_t)3 * k_new + 2 * (size_t)K + 8, 0);
  int32_t *rowsrc = map.data(), *colsrc = rowsrc + k_new, *jposn = colsrc + k_new, *etasrc = jposn + k_new, *wr = etasrc + K;
  for (int j = 0; j < K; ++j) {
    const int32_t p = e->h_eta_pos[(size_t)j];
    if (e->h_pos_core[(size_t)p] >= 0) { etasrc[j] = e->h_pos_core[(size_t)p]; continue; }
    int q = q_of(p);
    if (q < 0) {  // the FIRST eta at a position tells which variable the factorized basis held there
      const int64_t v = e->h_eta_leave[(size_t)j];
      if (v < e->ng) { set_err("refresh: basis bookkeeping inconsistent (structural variable at a slack position)"); return MLP_INVALID; }
      q = (int)qpos.size();
      qpos.push_back(p);
      wrow.push_back((int32_t)(v - e->ng));
    }
    etasrc[j] = -1 - q;
  }
  for (int t = 0; t < k_new; ++t) {
    const int32_t p = jpos[(size_t)t];
    jposn[t] = p;
    if (e->h_pos_core[(size_t)p] >= 0) { rowsrc[t] = e->h_pos_core[(size_t)p]; continue; }
    const int q = q_of(p);
    if (q < 0) { set_err("refresh: basis bookkeeping inconsistent (new core column without an eta)"); return MLP_INVALID; }
    rowsrc[t] = -1 - q;
  }
  for (int c = 0; c < k_new; ++c) {
    const int32_t r = R[(size_t)c];
    if (e->h_row_core[(size_t)r] >= 0) { colsrc[c] = e->h_row_core[(size_t)r]; continue; }
    int32_t p = e->h_rowcover_f[(size_t)r];  // position of the row's slack in the FACTORIZED basis: incremental_sets may have
    for (size_t q = 0; q < e->h_rc_old_rows.size(); ++q)  // moved h_rowcover_f on to the new basis already
      if (e->h_rc_old_rows[q] == r) { p = e->h_rc_old_vals[q]; break; }
    if (p < 0) { set_err("refresh: basis bookkeeping inconsistent (new core row without a basic slack)"); return MLP_INVALID; }
    colsrc[c] = -1 - p;
  }
  const int nq = (int)wrow.size();
  for (int q = 0; q < nq; ++q) wr[q] = wrow[(size_t)q];
  ST(stage_put(e, e->rf_map, map.data(), (size_t)3 * k_new + K + nq));
  const int32_t *d_rowsrc = e->rf_map, *d_colsrc = d_rowsrc + k_new, *d_jposn = d_colsrc + k_new, *d_etasrc = d_jposn + k_new,
                *d_wrow = d_etasrc + K;
  if (nq > 0)
    LAUNCH(e, k_rf_w, dim3(cdiv(k_old, 256), (unsigned)nq), 256, 0, e->dcsr_ptr, e->dcsr_idx, e->dcsr_val, d_wrow, k_old, e->Cinv, ld, e->rf_W, ld);
  LAUNCH(e, k_rf_t, cdiv(k_new, 32), 256, 0, e->Ginv, e->Kcap, K, d_etasrc, e->etaR, d_colsrc, k_new, e->Cinv, ld, e->rf_W, ld, e->rf_T);
  double* Cn = e->LUc;
  LAUNCH(e, k_rf_x0, dim3(cdiv(k_new, 256), (unsigned)std::min(k_new, 16384)), 256, 0, d_rowsrc, d_jposn, d_colsrc, k_new, e->Cinv, ld, e->rf_W, ld, Cn);
  LAUNCH(e, k_rf_ep, dim3(cdiv(k_new, 256), (unsigned)K), 256, 0, e->E, e->mld, d_jposn, k_new, e->rf_Ep, ld);
  for (int j0 = 0; j0 < K; j0 += GB_K)
    LAUNCH(e, k_gemm_sub<true>, dim3(cdiv(k_new, GB_T), cdiv(k_new, GB_T)), 256, 0, k_new, k_new, std::min(GB_K, K - j0),
           e->rf_Ep + (size_t)j0 * ld, ld, e->rf_T + j0, (int64_t)RF_CAP, Cn, ld);
  std::swap(e->Cinv, e->LUc);
  e->cnt.refreshes += 1;
  return MLP_OK;
}

// BasisSolver::reset (solver.rs:1286-1303) for B = [D | E_S], see DESIGN.md §4.  allow_refresh: the caller (mlp_pivot) has
// pushed the eta of the pivot that triggers the refactorization, so the eta file describes the whole change of the basis
// since the factors were made and may be folded into C^-1 instead of factorizing (refresh_inverse above).
static mlp_status refactor_impl(mlp_engine* e, bool allow_refresh = false) {
  const int64_t m = e->m, ng = e->ng;
  refac_stage(e, allow_refresh || e->refac_in_pivot ? "pivot: read-back, enter" : nullptr);
  e->refac_in_pivot = false;
  std::vector<int32_t> jpos, jslot, jvar, rowcover, R, prow, pval, gone;
  // Sparse storage, every basis change since the last refactorization on record: the new sets follow from the old ones.
  const bool incremental = e->sparse && e->chg_complete && (int64_t)e->h_pos_core.size() == m && (int64_t)e->h_rowcover_f.size() == m &&
                           (int64_t)e->h_Jpos_f.size() == e->k && (int64_t)e->h_R_sorted.size() == e->k;
  if (incremental) {
    ST(incremental_sets(e, jpos, R, prow, pval, gone));
    for (int32_t p : jpos) { jvar.push_back((int32_t)e->h_bvar[(size_t)p]); jslot.push_back(e->h_slot_of_row[(size_t)p]); }
  } else {
    rowcover.assign((size_t)m, -1);
    for (int64_t p = 0; p < m; ++p) {
      const int64_t v = e->h_bvar[p];
      if (v < ng) {
        jpos.push_back((int32_t)p);
        jvar.push_back((int32_t)v);
        if (!e->sparse && e->h_slot_of_row[p] < 0) { set_err("refactor: basic structural column missing from the cache"); return MLP_INVALID; }
        jslot.push_back(e->h_slot_of_row[p]);
      } else rowcover[v - ng] = (int32_t)p;
    }
    for (int64_t i = 0; i < m; ++i) if (rowcover[i] < 0) R.push_back((int32_t)i);
  }
  const int64_t k = (int64_t)jpos.size();
  if (!incremental && e->sparse && k > 1) {
    // order_simple (ordering.rs:4-21): columns by ascending entry count, FIFO — i.e. ascending basis position — within a
    // count.  (For a dense A every column has m entries and the order is the basis-position order built above.)
    std::vector<int32_t> ord((size_t)k);
    for (int64_t t = 0; t < k; ++t) ord[(size_t)t] = (int32_t)t;
    auto cnt = [&](int32_t t) { return e->h_csc_ptr[(size_t)jvar[(size_t)t] + 1] - e->h_csc_ptr[(size_t)jvar[(size_t)t]]; };
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return cnt(a) < cnt(b); });
    std::vector<int32_t> p2((size_t)k), v2((size_t)k), s2((size_t)k);
    for (int64_t t = 0; t < k; ++t) { p2[(size_t)t] = jpos[(size_t)ord[(size_t)t]]; v2[(size_t)t] = jvar[(size_t)ord[(size_t)t]]; s2[(size_t)t] = jslot[(size_t)ord[(size_t)t]]; }
    jpos.swap(p2); jvar.swap(v2); jslot.swap(s2);
  }
  if ((int64_t)R.size() != k) { set_err("refactor: basis bookkeeping inconsistent"); return MLP_INVALID; }
  std::vector<int32_t> Rsorted(R);  // R itself is overwritten with the factors' row order after a true factorization
  refac_stage(e, "host: index sets + column order");
  for (int32_t sl : e->h_pending_free) e->h_free_slots.push_back(sl);
  e->h_pending_free.clear();
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  e->spec_var = -1;
  e->ftran_var = -1;
  refac_stage(e, "drain both lanes");
  {
    size_t segs = 0;
    if (e->sparse) for (int32_t v : jvar) segs += (size_t)(e->h_col_seg[(size_t)v + 1] - e->h_col_seg[(size_t)v]);
    ST(stage_begin(e, 2 * (size_t)m + 10 * (size_t)k + segs + 2 * (size_t)e->K + 4 * RF_MAXK + 4 * prow.size() + 256));
  }
  bool refreshed = false;
  int64_t rf_core_before = 0;
  if (allow_refresh && k > 0 && k <= e->kcap && can_refresh(e)) {
    ST(refresh_inverse(e, jpos, R));
    refreshed = true;
    refac_stage(e, "refresh: C^-1 from the eta file");
  }
  ST(ensure_lu_capacity(e, k));
  // eta arena: the reference allows eta nnz up to lu nnz (solver.rs:1096-1097) ~ (k+1) dense columns
  {
    // The arena is dense, m doubles per eta, and bounded at 16 GB — a full arena just forces the next refactorization.  Dense
    // A: lu nnz ~ m k and an eta has m entries, so the file holds up to ~k etas: reserve 2k + 32.  Sparse A: the file is
    // short (lu nnz / nnz(alpha): tens to hundreds of etas) — reserving 2k + 32 columns would re-allocate gigabytes every
    // time k doubles (measured: 0.8 s per growth with peer mappings in place); follow the file's own length instead.
    const int64_t by_mem = std::max<int64_t>(1024, ((int64_t)16 << 30) / (8 * e->mld));
    const int64_t want = e->sparse ? std::min<int64_t>(2 * k + 32, 4 * e->K + 128) : 2 * k + 32;
    ST(ensure_eta_capacity(e, std::min<int64_t>(want, by_mem)));
  }
  refac_stage(e, "capacity (LU, eta arena)");
  e->k = k;
  e->K = 0;
  CU(cudaMemsetAsync(e->d_res->flags + 1, 0, sizeof(int), e->stream));
  CU(cudaMemsetAsync(e->etaLast, 0xff, (size_t)e->mld * sizeof(int32_t), e->stream));  // eta file is empty: no chains
  CU(cudaMemsetAsync(e->touched, 0, (size_t)e->mld, e->stream));
  // device maps by patches (the touched slack rows fit the scratch behind the refresh's maps) or in full
  const bool patch_maps = incremental && prow.size() <= (size_t)RF_MAXK && e->rf_map != nullptr;
  if (incremental) for (int32_t p : e->h_eta_pos) e->h_last_eta_of_row[(size_t)p] = -1;
  else std::fill(e->h_last_eta_of_row.begin(), e->h_last_eta_of_row.end(), -1);
  if (patch_maps) {
    int32_t* d_patch = e->rf_map + 3 * e->kcap;  // 2 RF_MAXK entries; stream-ordered behind the refresh kernels that read this area
    if (!prow.empty()) {  // rowcover: only the slack rows the basis changes touched
      ST(stage_put(e, d_patch, prow.data(), prow.size()));
      ST(stage_put(e, d_patch + RF_MAXK, pval.data(), pval.size()));
      LAUNCH(e, k_patch_i32, cdiv((int64_t)prow.size(), 256), 256, 0, e->rowcover, (const int32_t*)d_patch, (const int32_t*)(d_patch + RF_MAXK), (int)prow.size());
    }
    if (!gone.empty()) {  // rowcore: rows that left the core (a subset of the touched rows); the rows of the new core are set below
      std::vector<int32_t> minus((size_t)gone.size(), -1);
      ST(stage_put(e, d_patch, gone.data(), gone.size()));
      ST(stage_put(e, d_patch + RF_MAXK, minus.data(), minus.size()));
      LAUNCH(e, k_patch_i32, cdiv((int64_t)gone.size(), 256), 256, 0, e->rowcore, (const int32_t*)d_patch, (const int32_t*)(d_patch + RF_MAXK), (int)gone.size());
    }
  } else if (incremental) {
    ST(stage_put(e, e->rowcover, e->h_rowcover_f.data(), (size_t)m));
  } else {
    ST(stage_put(e, e->rowcover, rowcover.data(), (size_t)m));
  }
  if (e->sparse && !patch_maps) {  // also for an empty core: later refactorizations patch this map
    std::vector<int32_t> rowcore((size_t)m, -1);
    for (int64_t i = 0; i < k; ++i) rowcore[R[i]] = (int32_t)i;
    ST(stage_put(e, e->rowcore, rowcore.data(), (size_t)m));
  }
  if (k > 0) {
    ST(stage_put(e, e->Jpos, jpos.data(), (size_t)k));
    ST(stage_put(e, e->Jslot, jslot.data(), (size_t)k));
    ST(stage_put(e, e->Rp, R.data(), (size_t)k));
    if (e->sparse) {
      if (e->corevar_k > 0) LAUNCH(e, k_set_corepos, cdiv(e->corevar_k, 256), 256, 0, e->corepos, e->corevar, (int)e->corevar_k, 1);
      ST(stage_put(e, e->corevar, jvar.data(), (size_t)k));
      LAUNCH(e, k_set_corepos, cdiv(k, 256), 256, 0, e->corepos, e->corevar, (int)k, 0);
      e->corevar_k = k;
      if (patch_maps) LAUNCH(e, k_set_rowcore, cdiv(k, 256), 256, 0, e->rowcore, (const int32_t*)e->Rp, (int)k);
      // the core's segments
      std::vector<int32_t> cid, cfirst((size_t)k + 1, 0);
      for (int64_t t = 0; t < k; ++t) {
        cfirst[t] = (int32_t)cid.size();
        for (int64_t sg = e->h_col_seg[jvar[t]]; sg < e->h_col_seg[jvar[t] + 1]; ++sg) cid.push_back((int32_t)sg);
      }
      cfirst[k] = (int32_t)cid.size();
      e->ncseg = (int64_t)cid.size();
      if (e->ncseg > e->cseg_cap) {
        PoolScope pool(e->use_pool ? e->stream : nullptr);
        dev_free(e->cseg_id); dev_free(e->csum[0]); dev_free(e->csum[1]);
        e->cseg_cap = std::max<int64_t>(4 * e->ncseg, 1 << 15);
        ST(dev_alloc(&e->cseg_id, e->cseg_cap)); ST(dev_alloc(&e->csum[0], e->cseg_cap)); ST(dev_alloc(&e->csum[1], e->cseg_cap));
      }
      ST(stage_put(e, e->cseg_id, cid.data(), cid.size()));
      ST(stage_put(e, e->cseg_first, cfirst.data(), cfirst.size()));
    }
    refac_stage(e, "uploads: index maps, core segments");
    if (e->sparse) ST(build_core_rows(e, jvar));
    refac_stage(e, "compact core rows (DCSR)");
    if (refreshed) {
      // C^-1 is already the new core's.  Probe it against the new core (max |C C^-1 - I| over sampled columns) and count the
      // core's entries for the estimate of LUFactors::nnz below — one read-back; a failed probe falls through to the true
      // factorization (LUc, the old inverse's buffer, is scratch again).
      double r;
      ST(probe_inverse(e, k, &r, &rf_core_before));
      if (r <= e->rf_tol) e->rf_worst = std::max(e->rf_worst, r);
      if (!(r <= e->rf_tol)) { refreshed = false; e->cnt.refresh_rejects += 1; e->cnt.refreshes -= 1; }
    refac_stage(e, "refresh: accuracy probe + read back");
    }
    if (!refreshed) {
      if (e->sparse) {
        CU(cudaMemsetAsync(e->LUc, 0, (size_t)e->kcap * k * sizeof(double), e->stream));
        LAUNCH(e, k_extract_core_seg, cdiv(e->ncseg, 8), 256, 0, e->csc_ptr, e->csc_idx, e->csc_val, e->seg_col, e->seg_off, e->cseg_id,
               (int)e->ncseg, e->corepos, e->rowcore, e->LUc, e->kcap);
        CU(cudaMemsetAsync(e->d_nnzcnt, 0, 2 * sizeof(unsigned long long), e->stream));
        LAUNCH(e, k_core_row_counts, cdiv(k, 256), 256, 0, e->LUc, e->kcap, (int)k, e->lu_rcnt, e->d_nnzcnt);
      } else
        LAUNCH(e, k_extract_core, dim3(cdiv(k, 256), (unsigned)k), 256, 0, e->Bcols, e->mld, (int)k, e->Rp, e->Jslot, e->LUc, e->kcap);
      refac_stage(e, "extract core + row counts");
      int* flags = e->d_res->flags;
      for (int j0 = 0; j0 < (int)k;) {
        const int rows = (int)k - j0;
        // widest panel whose rows x nb block (+ row ids, permutation) fits in shared memory; else work in place in global memory
        int nb = LU_NB, use_smem = 0;
        const size_t per_row = e->sparse ? 12 : 8;  // row ids + permutation (+ row entry counts)
        for (int cand = LU_NB; cand >= 4; cand /= 2)
          if ((size_t)rows * cand * 8 + (size_t)rows * per_row <= e->smem_optin) { nb = cand; use_smem = 1; break; }
        nb = std::min(nb, rows);
        const size_t smem = use_smem ? (size_t)rows * nb * 8 + (size_t)rows * per_row : 0;
        const int pt = std::max(64, std::min(1024, (rows + 31) / 32 * 32));  // one row per thread
        LAUNCH(e, k_lu_panel, 1, pt, smem, e->LUc, e->kcap, (int)k, j0, nb, e->Rp, e->sparse ? e->lu_rcnt : (int32_t*)nullptr, flags,
               e->lu_aff, e->lu_aff + 64, e->lu_aff + 128, e->lu_perm, use_smem);
        if ((int)k > nb) LAUNCH(e, k_lu_swap_solve, cdiv(k - nb, 8), 256, 0, e->LUc, e->kcap, (int)k, j0, nb, e->lu_aff, e->lu_aff + 64,
                                e->lu_aff + 128, flags);
        const int rem = rows - nb;
        if (rem > 0) LAUNCH(e, k_lu_trailing, dim3(cdiv(rem, LU_NC), cdiv(rem, 256)), 256, 0, e->LUc, e->kcap, (int)k, j0, nb, flags);
        j0 += nb;
      }
      refac_stage(e, "LU panels / swap-solve / trailing");
      {  // (L U)^-1, one CTA per column
        const size_t need = (size_t)k * sizeof(double);
        const int use_smem = need <= e->smem_optin ? 1 : 0;
        if (k >= e->inv_blocked_min) {
          // blocked substitution on all columns at once (dense_block.cuh): X = I; forward through L, backward through U
          const int nbk = cdiv(k, 32);
          LAUNCH(e, k_set_identity, dim3(cdiv(k, 256), (unsigned)k), 256, 0, e->Cinv, e->kcap, (int)k);
          for (int b = 0; b < nbk; ++b) {  // L y = e: X stays lower triangular, only columns < (b+1)*32 are non-zero
            const int r0 = b * 32, nb = std::min<int>(32, (int)k - r0), nc = std::min<int>((int)k, r0 + nb), below = (int)k - (r0 + nb);
            LAUNCH(e, k_tri_block<true>, cdiv(nc, 128), 128, 0, e->LUc, e->kcap, r0, nb, e->Cinv, e->kcap, nc);
            if (below > 0)
              LAUNCH(e, k_gemm_sub<true>, dim3(cdiv(below, GB_T), cdiv(nc, GB_T)), 256, 0, below, nc, nb, e->LUc + (size_t)r0 * e->kcap + r0 + nb,
                     e->kcap, e->Cinv + r0, e->kcap, e->Cinv + r0 + nb, e->kcap);
          }
          for (int b = nbk - 1; b >= 0; --b) {  // U x = y over all k columns
            const int r0 = b * 32, nb = std::min<int>(32, (int)k - r0);
            LAUNCH(e, k_tri_block<false>, cdiv(k, 128), 128, 0, e->LUc, e->kcap, r0, nb, e->Cinv, e->kcap, (int)k);
            if (r0 > 0)
              LAUNCH(e, k_gemm_sub<true>, dim3(cdiv(r0, GB_T), cdiv(k, GB_T)), 256, 0, r0, (int)k, nb, e->LUc + (size_t)r0 * e->kcap, e->kcap,
                     e->Cinv + r0, e->kcap, e->Cinv, e->kcap);
          }
        } else if (k <= 256) LAUNCH(e, k_core_inverse_pf<1>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
        else if (k <= 512) LAUNCH(e, k_core_inverse_pf<2>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
        else if (k <= 1024) LAUNCH(e, k_core_inverse_pf<4>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
        else if (k <= 2048) LAUNCH(e, k_core_inverse_pf<8>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
        else if (k <= 4096) LAUNCH(e, k_core_inverse_pf<16>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
        else LAUNCH(e, k_core_inverse, (unsigned)k, 256, use_smem ? need : 0, e->LUc, e->kcap, (int)k, e->Cinv, flags, use_smem);
      }
      refac_stage(e, "explicit inverse");
      if (e->sparse) {
        LAUNCH(e, k_count_offdiag, dim3(cdiv(k, 256), cdiv(k, 64)), 256, 0, e->LUc, e->kcap, (int)k, e->d_nnzcnt + 1);
        CU(cudaMemcpyAsync(&e->d_res->i[4], e->d_nnzcnt, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, e->stream));
      }
      ST(fetch_res(e, e->lane[0]));
      if (e->h_res->flags[1]) { set_err("singular basis"); return MLP_SINGULAR; }
      // the factorization permuted the core's rows (Rp): column c of C^-1 belongs to row Rp[c] — the next refresh needs that order
      if (e->sparse) ST(d2h(e, R.data(), e->Rp, (size_t)k * sizeof(int32_t)));
      if (e->sparse && e->refac_trace) {  // calibration of the refresh probe: the same measure on a freshly factorized inverse
        double r;
        int64_t ce;
        ST(probe_inverse(e, k, &r, &ce));
        e->rf_worst_true = std::max(e->rf_worst_true, r);
      }
    }
    refac_stage(e, "count off-diagonal + read back");
  } else {
    if (e->sparse) ST(build_core_rows(e, jvar));  // empty
    if (e->sparse && e->corevar_k > 0) {
      LAUNCH(e, k_set_corepos, cdiv(e->corevar_k, 256), 256, 0, e->corepos, e->corevar, (int)e->corevar_k, 1);
      e->corevar_k = 0;
    }
    CU(cudaStreamSynchronize(e->stream));
  }
  // LUFactors::nnz (lu.rs:52-54): lower.nondiag + upper.nondiag + m.  The entries of the k structural basic columns in
  // slack-covered rows go to U unchanged; the core contributes the off-diagonal entries of its factors, FILL-IN INCLUDED
  // (counted on the device; exact zeros are not stored, lu.rs:253-255).  Dense A: no zeros, k(k-1) + (m-k)k + m in closed
  // form.  The refactor rule (solver.rs:1096-1097) is the reference's, applied to the factors the engine really has: with
  // the same column order and pivot rule (ties aside) their size tracks the reference's.
  if (e->sparse) {
    int64_t nz = 0;
    for (int32_t v : jvar) nz += e->h_csc_ptr[(size_t)v + 1] - e->h_csc_ptr[(size_t)v];
    if (refreshed) {
      // no factors to count: the part outside the core is exact, the core's L\U is taken to fill as it did at the last true
      // factorization (off-diagonal entries of the factors per entry of the core)
      e->lu_nnz = (nz - rf_core_before) + (int64_t)((double)rf_core_before * e->fill_true) + m;
    } else {
      const int64_t core_before = k > 0 ? e->h_res->i[4] : 0, core_offdiag = k > 0 ? e->h_res->i[5] : 0;
      e->lu_nnz = (nz - core_before) + core_offdiag + m;
      e->fill_true = core_before > 0 ? (double)core_offdiag / (double)core_before : 1.0;
      e->pivots_since_lu = 0;
    }
  } else e->lu_nnz = k * (k - 1) + (m - k) * k + m;
  if (e->sparse) {  // the sets of the factorized basis, for the next refresh / the next incremental set-up
    if (incremental) {
      for (int32_t p : e->h_Jpos_f) e->h_pos_core[(size_t)p] = -1;
      for (int32_t r : e->h_R_f) e->h_row_core[(size_t)r] = -1;
    } else {
      e->h_pos_core.assign((size_t)m, -1);
      e->h_row_core.assign((size_t)m, -1);
      e->h_rowcover_f.swap(rowcover);
    }
    for (int64_t t = 0; t < k; ++t) { e->h_pos_core[(size_t)jpos[(size_t)t]] = (int32_t)t; e->h_row_core[(size_t)R[(size_t)t]] = (int32_t)t; }
    e->h_Jpos_f.swap(jpos);
    e->h_R_f.swap(R);          // the factors' row order (permuted by a true factorization)
    e->h_R_sorted.swap(Rsorted);
    e->h_eta_pos.clear();
    e->h_eta_leave.clear();
    e->h_rc_old_rows.clear();
    e->h_rc_old_vals.clear();
    e->chg_complete = true;
    e->inv_valid = true;
  }
  ST(stage_end(e));
  e->cnt.refactors += 1;
  e->cnt.k_structural = k;
  e->refac_k_sum += (double)k;
  refac_stage(e, "host: lu nnz");
  return MLP_OK;
}

