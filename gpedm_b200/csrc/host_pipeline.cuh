// libgpedm_b200 host orchestration - launch sequence of one evaluation: tile launches, blocked Cholesky with look-ahead, recursive triangular inverse, graph replay.
// Part of the single translation unit gpedm_core.cu (included in order; not a standalone header).
#pragma once

// ------------------------------------------------------------------------------- tile launches
// layout: 0 = (MN, MN), 1 = (MN, K), 2 = (K, K) operand layouts (see dgemm_tile.cuh).
static int launch_tiles(gpedm_fit_s* h, cudaStream_t st, const TileOp& op_in, int layout, bool split = false) {
  if (op_in.ntiles <= 0) return 0;
  TileOp op = op_in;
  op.tag = st == h->aux ? 1 : (st == h->fill ? 2 : 0);
  if (split && layout == 0)  // latency-critical launch with fewer tiles than SMs: 4 row strips per tile
    tile_once_kernel<false, false, GemmCfgSplit>
        <<<op.ntiles * (TB / GemmCfgSplit::TM), GemmCfgSplit::THREADS, GemmCfgSplit::SMEM_BYTES, st>>>(op);
  else if (!g_no_half_tiles && op.type != OP_PANEL && op.ntiles * 2 <= h->num_sms) {
    // launches with fewer tiles than half the SMs (low levels of the triangular inverse, small N):
    // 128 x 64 half tiles double the number of busy SMs.  (Measured slower than full tiles once the
    // grid fills the chip, so only used below that point.  The panel stays on row strips / full
    // tiles: it is updated in place and a column half would read columns another CTA overwrites.)
    const int grid = op.ntiles * 2;
    if (layout == 0)
      tile_once_kernel<false, false, GemmCfgHalf><<<grid, GemmCfgHalf::THREADS, GemmCfgHalf::SMEM_BYTES, st>>>(op);
    else if (layout == 1)
      tile_once_kernel<false, true, GemmCfgHalf><<<grid, GemmCfgHalf::THREADS, GemmCfgHalf::SMEM_BYTES, st>>>(op);
    else
      tile_once_kernel<true, true, GemmCfgHalf><<<grid, GemmCfgHalf::THREADS, GemmCfgHalf::SMEM_BYTES, st>>>(op);
  } else if (layout == 0)
    tile_once_kernel<false, false><<<op.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(op);
  else if (layout == 1)
    tile_once_kernel<false, true><<<op.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(op);
  else
    tile_once_kernel<true, true><<<op.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(op);
  h->launches++;
  return 0;
}

static TileOp make_op(int type, int ntiles, int nb) {
  TileOp op;
  memset(&op, 0, sizeof(op));
  op.type = type;
  op.ntiles = ntiles;
  op.nb = nb;
  return op;
}


// Triangular inverse X = L^-1 by recursive doubling, restricted to what is computable once the
// first `ncols` block columns of L are final.  Level s (block size in tiles), pair p (tile rows/cols
// [2ps, 2ps+2s)): X21 = -X22 (L21 X11) in two launches,
//   W = L21 X11   needs X11 (level s/2, pair 2p) and block columns < 2ps+s of L,
//   X21 = -X22 W  needs X22 (level s/2, pair 2p+1), i.e. block columns < 2ps+2s.
// h->trtri_w_done / trtri_x_done count the full pairs already launched per level, so repeated calls
// with growing `ncols` launch every product exactly once and in dependency order (ascending s within a
// call, calls in stream order).  In particular the W product of the TOP level - a quarter of all
// the flops of the inverse - becomes available when the Cholesky is only half way.
// `final` also launches the ragged last pair of each level.
static int enqueue_trtri_levels(gpedm_fit_s* h, cudaStream_t st, int ncols, bool final) {
  const int nb = h->nb;
  const size_t ld = h->Np;
  int lv = 0;
  for (int s = 1; s < nb; s *= 2, ++lv) {
    const int pf = nb / (2 * s);  // full pairs of this level
    const int rem = nb - 2 * s * pf;
    const bool ragged = final && rem > s;
    // ---- W phase
    int wready = h->trtri_w_done[lv];
    while (wready < pf && 2 * wready * s + s <= ncols && (lv == 0 || h->trtri_x_done[lv - 1] >= 2 * wready + 1)) ++wready;
    {
      const int p0 = h->trtri_w_done[lv], nfull = wready - p0;
      const int ntl = nfull * s * s + (ragged ? (rem - s) * s : 0);
      if (ntl > 0) {
        TileOp w = make_op(OP_TRTRI_W, ntl, nb);
        w.s = s;
        w.p0 = p0;
        w.nfull = nfull;
        w.P0 = h->bufA;
        w.P1 = h->bufB;
        w.P2 = h->bufC;
        w.ld0 = w.ld1 = w.ld2 = ld;
        int rc = launch_tiles(h, st, w, 1);
        if (rc) return rc;
        h->trtri_w_done[lv] = wready;
      }
    }
    // ---- X phase
    {
      const int xready = std::min(h->trtri_w_done[lv], std::min(pf, ncols / (2 * s)));
      const int p0 = h->trtri_x_done[lv], nfull = std::max(0, xready - p0);
      const int ntl = nfull * s * s + (ragged ? (rem - s) * s : 0);
      if (ntl > 0) {
        TileOp x = make_op(OP_TRTRI_X, ntl, nb);
        x.s = s;
        x.p0 = p0;
        x.nfull = nfull;
        x.P0 = h->bufB;
        x.P1 = h->bufC;
        x.P2 = h->bufB;
        x.ld0 = x.ld1 = x.ld2 = ld;
        int rc = launch_tiles(h, st, x, 1);
        if (rc) return rc;
        h->trtri_x_done[lv] = p0 + nfull;
      }
    }
  }
  return 0;
}

// Phase-timing events must stay usable from outside when the launch sequence is captured into a
// CUDA graph: record them as "external" event nodes during capture.
static thread_local bool g_capturing = false;  // set only inside enqueue_eval while capturing
static cudaError_t record_phase_event(cudaEvent_t ev, cudaStream_t st) {
  return g_capturing ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
}

// ------------------------------------------------------------------------------- pipeline
// Enqueue the device work of one evaluation on h->stream.  h->hparams must be filled.
// stages: through z/a (need_inverse_products=false stops after z for the sequential helper).
static int enqueue_eval_raw(gpedm_fit_s* h, bool full, bool through_z_only) {
  cudaStream_t st = h->stream;
  const int nb = h->nb, Np = h->Np, N = (int)h->N, d = h->d;
  const size_t ld = Np;
  NvtxRange nvtx_phase("gpedm:assemble");
  CU(cudaMemcpyAsync(h->dparams, h->hparams, sizeof(EvalParams), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(h->dinfo, 0, 4, st));
  CU(record_phase_event(h->ev[0], st));
  // --- assemble Sigma (lower tiles) into bufA (or take the caller's matrix: gpedm_covinv)
  if (h->dsigma_in) {
    load_sigma_kernel<<<h->ntiles, 256, 0, st>>>(h->bufA, h->bufD, ld, N, h->dsigma_in);
    h->launches++;
  } else {
    int smem = 2 * d * TB * 8 + 2 * TB * 4;
    assemble_sigma_kernel<<<h->ntiles * h->tile_split, 256, smem, st>>>(h->bufA, h->bufD, ld, N, h->dX, h->dpop, h->drhotab,
                                                                          h->dparams, h->tile_split);
    h->launches++;
  }
  CU(record_phase_event(h->ev[1], st));
  NvtxRange::next("gpedm:potrf");
  if (!g_no_dag) {
    // --- Cholesky + triangular inverse: ONE persistent, dependency-driven launch (dag_factor.cuh)
    CU(cudaMemsetAsync(h->dsync, 0, (size_t)(1 + DAG_NFLAG * nb * nb) * 4, st));
    DagArgs da;
    da.fits = h->dfit;
    da.tasks = through_z_only ? h->dtasks + h->factor_off : h->dtasks;
    da.ntasks = through_z_only ? h->ntasks_factor : h->ntasks;
    da.counter = h->dsync;
    da.strips = g_dag_strips ? 1 : 0;
    da.trsm = g_dag_trsm;
    dag_factor_kernel<<<std::min(h->num_sms, da.ntasks), DAG_THREADS, DAG_SMEM_BYTES, st>>>(da);
    h->launches++;
    CU(record_phase_event(h->ev[2], st));
    NvtxRange::next("gpedm:trtri");
    CU(record_phase_event(h->ev[3], st));
  } else {
  // --- Cholesky, right-looking over 128-wide block columns, with look-ahead and grouped updates.
  // Steps are processed in groups [g0, g1) of G block columns (G = GPEDM_POTRF_GROUP while the
  // trailing matrix is large, 1 near the end); [g1, g2) is the next group.
  //   aux (high priority): for q in [g0, g1):
  //        column q <- pending updates of the group's earlier steps (rank 128 (q-g0))
  //        diag(q) -> panel(q)
  //      then columns [g1, g2) <- all G steps of the group (rank 128 G), so the next group's whole
  //      chain depends only on this stream;
  //   main: columns >= g2 <- all G steps in ONE rank-128G update: each trailing tile is read and
  //        written once per group instead of once per step and the k loop is G times longer.
  // The chain of group g+1 overlaps the bulk update of group g.
  {
    cudaStream_t ax = g_no_lookahead ? st : h->aux;
    CU(cudaEventRecord(h->evFork, st));
    CU(cudaStreamWaitEvent(ax, h->evFork, 0));
    auto trail = [&](cudaStream_t s_, int k0, int kc, int jlo, int ntl) -> int {
      TileOp t = make_op(OP_TRAIL, ntl, nb);
      t.k = k0;
      t.kcount = kc;
      t.jlo = jlo;
      t.P2 = h->bufA;
      t.ld2 = ld;
      return launch_tiles(h, s_, t, 0, s_ == ax && !g_no_split && ntl * 4 * 2 <= g_split_update2 * h->num_sms);
    };
    // group size: GPEDM_POTRF_GROUP if set, else 4 from 64 block columns up (bulk-dominated; measured
    // 20.46 -> 20.15 ms at N=8192 together with group_min 16 / fill_tail 40), 2 below (4.07 vs 4.17 ms at N=4096)
    const int gsz = g_group > 0 ? g_group : (nb >= 64 ? 4 : 2);
    auto group_size = [&](int start) { return (nb - start - 1 >= g_group_min) ? gsz : 1; };
    // Early triangular inverse: level s, pair p (tile rows/cols [2ps, 2ps+2s)) only needs block
    // columns < 2s(p+1) of L and the diagonal-block inverses, so it can run as soon as the Cholesky
    // has passed that column - on a low-priority stream, filling the SMs the chain-bound tail of the
    // factorisation leaves idle.  Checkpoints every 8 block columns.
    cudaStream_t fl = (g_no_lookahead || g_no_fill) ? nullptr : h->fill;
    if (fl) CU(cudaStreamWaitEvent(fl, h->evFork, 0));
    for (int lv = 0; lv < 16; ++lv) h->trtri_w_done[lv] = h->trtri_x_done[lv] = 0;
    int prev_bulk = -1;  // index of the event recorded after the previous group's bulk update
    int g0 = 0, G = group_size(0);
    while (g0 < nb) {
      const int g1 = std::min(nb, g0 + G);
      for (int q = g0; q < g1; ++q) {
        if (q > g0) {
          int rc = trail(ax, g0, q - g0, q, nb - q);  // column q only
          if (rc) return rc;
        }
        if (g_diag_ref)
          diag_potf2_inv_kernel<<<1, 256, DIAG_SMEM_BYTES, ax>>>(h->bufA + (size_t)q * TB * (ld + 1), ld,
                                                                  h->bufB + (size_t)q * TB * (ld + 1), ld, q, N,
                                                                  h->dlogdet, h->dinfo);
        else
          diag_block_kernel<<<1, DG_THREADS, DG_SMEM_BYTES, ax>>>(h->bufA + (size_t)q * TB * (ld + 1), ld,
                                                                  h->bufB + (size_t)q * TB * (ld + 1), ld, q, N,
                                                                  h->dlogdet, h->dinfo);
        h->launches++;
        if (nb - q - 1 > 0) {
          TileOp pan = make_op(OP_PANEL, nb - q - 1, nb);
          pan.k = q;
          pan.P1 = h->bufB;
          pan.P2 = h->bufA;
          pan.ld1 = pan.ld2 = ld;
          int rc = launch_tiles(h, ax, pan, 0, !g_no_split && pan.ntiles * 4 * 2 <= g_split_panel2 * h->num_sms);
          if (rc) return rc;
        }
      }
      if (fl && g1 % 8 == 0 && g1 < nb && nb - g1 <= g_fill_tail) {  // only once the factorisation is chain-bound
        CU(cudaEventRecord(h->evCol[g1 / 8], ax));
        CU(cudaStreamWaitEvent(fl, h->evCol[g1 / 8], 0));
        int rc = enqueue_trtri_levels(h, fl, g1, false);
        if (rc) return rc;
      }
      if (g1 >= nb) break;
      const int Gn = group_size(g1);
      const int g2 = std::min(nb, g1 + Gn);
      CU(cudaEventRecord(h->evPanel[g1 - 1], ax));
      if (prev_bulk >= 0) CU(cudaStreamWaitEvent(ax, h->evTrail[prev_bulk], 0));
      int next_cols_tiles = 0;
      for (int c = g1; c < g2; ++c) next_cols_tiles += nb - c;
      int rc = trail(ax, g0, g1 - g0, g1, next_cols_tiles);  // columns [g1, g2)
      if (rc) return rc;
      CU(cudaStreamWaitEvent(st, h->evPanel[g1 - 1], 0));
      if (g2 < nb) {
        const int m = nb - g2;
        rc = trail(st, g0, g1 - g0, g2, m * (m + 1) / 2);  // columns >= g2
        if (rc) return rc;
      }
      CU(cudaEventRecord(h->evTrail[g1 - 1], st));
      prev_bulk = g1 - 1;
      g0 = g1;
      G = Gn;
    }
    CU(cudaEventRecord(h->evJoin, ax));
    CU(cudaStreamWaitEvent(st, h->evJoin, 0));
    if (fl) {
      CU(cudaEventRecord(h->evFill, fl));
      CU(cudaStreamWaitEvent(st, h->evFill, 0));
    }
  }
  CU(record_phase_event(h->ev[2], st));
  NvtxRange::next("gpedm:trtri");
  // --- triangular inverse X = L^-1 by recursive doubling (diagonal blocks already in bufB): whatever the
  // fill stream has not done yet, level by level
  {
    int rc = enqueue_trtri_levels(h, st, nb, true);
    if (rc) return rc;
  }
  CU(record_phase_event(h->ev[3], st));
  }  // !g_no_dag
  NvtxRange::next("gpedm:vec");
  // --- z = X Y
  trmv_n_part_kernel<<<h->ntiles, 256, 0, st>>>(h->bufB, ld, h->dY, h->dpart, ld);
  reduce_parts_kernel<<<ceil_div(Np, 256), 256, 0, st>>>(h->dpart, ld, nb, 0, h->dz, Np);
  h->launches += 2;
  if (through_z_only) {
    CU(record_phase_event(h->ev[4], st));
    CU(cudaMemcpyAsync(h->hinfo, h->dinfo, 4, cudaMemcpyDeviceToHost, st));
    return 0;
  }
  // --- a = X^T z
  trmv_t_part_kernel<<<h->ntiles, 256, 0, st>>>(h->bufB, ld, h->dz, h->dpart, ld);
  reduce_parts_kernel<<<ceil_div(Np, 256), 256, 0, st>>>(h->dpart, ld, nb, 1, h->da, Np);
  h->launches += 2;
  CU(record_phase_event(h->ev[4], st));
  NvtxRange::next("gpedm:lauum");
  // --- S = Sigma^-1 = X^T X (lower tiles) into bufC (already done by the persistent kernel unless GPEDM_NO_DAG)
  if (g_no_dag) {
    TileOp lu = make_op(OP_LAUUM, h->ntiles, nb);
    lu.P0 = h->bufB;
    lu.P2 = h->bufC;
    lu.ld0 = lu.ld2 = ld;
    int rc = launch_tiles(h, st, lu, 2);
    if (rc) return rc;
  }
  CU(record_phase_event(h->ev[5], st));
  NvtxRange::next("gpedm:trace");
  // --- gradient trace terms + scalar reductions
  {
    const int DC = d <= 8 ? 8 : (d <= 12 ? 12 : (d <= 16 ? 16 : 32));
    const int smem = (2 * d * TB + 2 * TB + 8 * (DC + 3)) * 8 + 2 * TB * 4;
    const int cs = h->tile_split, grid = h->ntiles * cs;
#define GPEDM_TRACE(DCV)                                                                                         \
  grad_trace_kernel<DCV><<<grid, 256, smem, st>>>(h->bufC, h->bufD, ld, N, h->dX, h->dpop, h->da, h->dparams, \
                                                  h->dtracepart, cs)
    if (DC == 8)
      GPEDM_TRACE(8);
    else if (DC == 12)
      GPEDM_TRACE(12);
    else if (DC == 16)
      GPEDM_TRACE(16);
    else
      GPEDM_TRACE(32);
#undef GPEDM_TRACE
    h->launches++;
    if (full) {
      extract_diag_kernel<<<ceil_div(N, 256), 256, 0, st>>>(h->bufC, ld, N, h->ddiagS);
      h->launches++;
    }
    final_reduce_kernel<<<1, 256, 0, st>>>(h->dlogdet, nb, h->dz, h->da, h->ddiagS, N, h->dtracepart, h->ntiles * h->tile_split,
                                           d + 3, full ? 1 : 0, h->dscal);
    h->launches++;
  }
  CU(record_phase_event(h->ev[6], st));
  CU(cudaMemcpyAsync(h->hscal, h->dscal, NSCAL * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->hinfo, h->dinfo, 4, cudaMemcpyDeviceToHost, st));
  return 0;
}


// One evaluation = a fixed launch sequence on two streams (parameters live in device memory), so
// it is captured once per handle and variant into a CUDA graph and replayed: one graph launch
// instead of hundreds of kernel launches / event waits per evaluation.
static int enqueue_eval(gpedm_fit_s* h, bool full, bool through_z_only = false) {
  // Above ~48 block columns the bulk kernels dominate and the stream-priority look-ahead (which the
  // graph replay does not honour as well as live streams do, measured) matters more than launch cost.
  if (g_no_graph || h->nb > g_graph_max_nb) return enqueue_eval_raw(h, full, through_z_only);
  const int v = through_z_only ? 2 : (full ? 1 : 0);
  if (!h->graph_exec[v]) {
    const int64_t l0 = h->launches;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    g_capturing = true;
    int rc = enqueue_eval_raw(h, full, through_z_only);
    g_capturing = false;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph_exec[v], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return set_err(GPEDM_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e));
    h->graph_launches[v] = h->launches - l0;
    h->launches = l0;
  }
  CU(cudaGraphLaunch(h->graph_exec[v], h->stream));
  h->launches += h->graph_launches[v];
  return 0;
}

static int collect_phase_times(gpedm_fit_s* h) {
  // ev: 0 start, 1 after assemble, 2 after potrf, 3 after trtri, 4 after vec, 5 after lauum, 6 after trace
  float t;
  CU(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
  h->phase_ms[GPEDM_PHASE_ASSEMBLE] = t;
  CU(cudaEventElapsedTime(&t, h->ev[1], h->ev[2]));
  h->phase_ms[GPEDM_PHASE_POTRF] = t;
  CU(cudaEventElapsedTime(&t, h->ev[2], h->ev[3]));
  h->phase_ms[GPEDM_PHASE_TRTRI] = t;
  CU(cudaEventElapsedTime(&t, h->ev[3], h->ev[4]));
  h->phase_ms[GPEDM_PHASE_VEC] = t;
  CU(cudaEventElapsedTime(&t, h->ev[4], h->ev[5]));
  h->phase_ms[GPEDM_PHASE_LAUUM] = t;
  CU(cudaEventElapsedTime(&t, h->ev[5], h->ev[6]));
  h->phase_ms[GPEDM_PHASE_TRACE] = t;
  CU(cudaEventElapsedTime(&t, h->ev[0], h->ev[6]));
  h->phase_ms[GPEDM_PHASE_TOTAL] = t;
  return 0;
}
