// tribs_ref_binding.h — the binding a tRIBS maintainer adds to drive the B200 path from the
// reference's own objects. It is compiled against the reference headers (it is the body that
// tHydroModel::UnSaturatedZone / ResetGW / SaturatedZone / Reset forward to) and talks to
// libtribs_gpu.so only through the C ABI of include/tribs_gpu.h.
//
//   attach():  walks tMesh<tCNode> exactly like the reference's own loops
//              (tMeshListIter::FirstP/NextP/IsActive, tEdge::getCCWEdg ...) and flattens it into
//              tribs_mesh_t / tribs_params_t / tribs_state_t  -> tribs_gpu_init
//   UnSaturatedZone(dt) / ResetGW() / SaturatedZone(dtGW) / Reset():  tribs_gpu_step with the
//              flags the reference computes on the host (tRunTimer arithmetic, fmod tests)
//   attachET() / callEvapoPotential() / callEvapoTrans(flag):  the cell loops of tEvapoTrans and
//              tIntercept run as kernel (1); tEvapoTrans keeps what it does once per call
//              (SetEnvironment: calendar + sun geometry, the station records of the hour)
//   syncToNodes(mask):  tribs_gpu_sync + scatter into tCNode through its own setters, so
//              tKinemat (channel routing), tCOutput, tWaterBalance and tRestart keep working on
//              the host unchanged.
//
// In this repo the header is compiled into the reference test harness (mode --gpu), which is
// how tests/test_hybrid.py checks that the reference's Simulator can drive the GPU path and
// that the outlet hydrograph of tKinemat stays within 1e-6 of the pure reference run.
#pragma once
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>

#include "tribs_gpu.h"

struct TribsGpuBinding {
  // ---- C ABI entry points (resolved with dlsym so the host binary has no link-time dependency)
  void *lib = nullptr;
  decltype(&tribs_gpu_init) p_init = nullptr;
  decltype(&tribs_gpu_step) p_step = nullptr;
  decltype(&tribs_gpu_sync) p_sync = nullptr;
  decltype(&tribs_gpu_destroy) p_destroy = nullptr;
  decltype(&tribs_gpu_last_error) p_err = nullptr;
  tribs_handle_t h = nullptr;

  tMesh<tCNode> *mesh = nullptr;
  tRunTimer *timer = nullptr;
  std::vector<tCNode *> cells;                       // active cells in sweep (list) order
  std::unordered_map<tNode *, int> pos;
  std::vector<std::vector<double>> stage;            // host SoA staging, one vector per field
  std::vector<double> rain;

  [[noreturn]] void die(const char *what) {          // the reference's error convention
    std::cout << "\nError: tribs_gpu: " << what << ": " << (p_err ? p_err() : dlerror()) << std::endl;
    exit(2);
  }
  int idx(tNode *n) const { auto k = pos.find(n); return k == pos.end() ? -1 : k->second; }

  // ---- reach partitions (what tGraph / tParallel do in the MPI build). `rank_h` holds this process's ranks:
  // one handle per MPI rank in a tRIBSpar-style host (it passes its own rank to attach(), gathers the 64-byte
  // handles of tribs_gpu_ipc_export with MPI_Allgather where tParallel::initialize sets up MPI, and hands them
  // to tribs_gpu_attach_peers), or all ranks of the run when one process drives several GPUs (attach() with
  // nranks > 1 below: tribs_gpu_attach_local). `owner[i]` = reach2partition[tCNode::getReach()] of sweep cell i.
  std::vector<tribs_handle_t> rank_h;
  std::vector<int32_t> owner;
  decltype(&tribs_gpu_run) p_run = nullptr;
  decltype(&tribs_gpu_wait) p_wait = nullptr;
  decltype(&tribs_gpu_attach_local) p_attach_local = nullptr;
  decltype(&tribs_gpu_ipc_export) p_ipc_export = nullptr;
  decltype(&tribs_gpu_attach_peers) p_attach_peers = nullptr;

  void attach(tMesh<tCNode> *m, tKinemat *flow, tHydroModel *hydro, tRunTimer *t, const char *libpath,
              int nranks = 1, const std::vector<int32_t> *cell_owner = nullptr, int max_batch_steps = 64) {
    mesh = m; timer = t;
    lib = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!lib) die("dlopen");
    p_init = (decltype(p_init))dlsym(lib, "tribs_gpu_init");
    p_step = (decltype(p_step))dlsym(lib, "tribs_gpu_step");
    p_sync = (decltype(p_sync))dlsym(lib, "tribs_gpu_sync");
    p_destroy = (decltype(p_destroy))dlsym(lib, "tribs_gpu_destroy");
    p_err = (decltype(p_err))dlsym(lib, "tribs_gpu_last_error");
    if (!p_init || !p_step || !p_sync || !p_destroy || !p_err) die("dlsym");
    auto p_abi = (decltype(&tribs_gpu_abi_version))dlsym(lib, "tribs_gpu_abi_version");
    if (!p_abi || p_abi() != TRIBS_GPU_ABI_VERSION) {       // struct layouts of another header version
      std::fprintf(stderr, "\nError: tribs_gpu: %s was built for ABI %d, this binding for %d\n", libpath, p_abi ? p_abi() : -1,
                   TRIBS_GPU_ABI_VERSION);
      std::exit(2);
    }

    tMeshListIter<tCNode> it(mesh->getNodeList());
    for (tCNode *cn = it.FirstP(); it.IsActive(); cn = it.NextP()) { pos[(tNode *)cn] = (int)cells.size(); cells.push_back(cn); }
    const int n = (int)cells.size();
    auto col = [&](auto fn) { std::vector<double> v(n); for (int i = 0; i < n; i++) v[i] = fn(cells[i]); return v; };
    auto icol = [&](auto fn) { std::vector<int32_t> v(n); for (int i = 0; i < n; i++) v[i] = fn(cells[i]); return v; };

    // --- mesh
    auto boundary = icol([](tCNode *c) { return c->getBoundaryFlag(); });
    auto z = col([](tCNode *c) { return c->getZ(); });
    auto varea = col([](tCNode *c) { return c->getVArea(); });
    auto fslope = col([](tCNode *c) { return c->getFlowEdg()->getSlope(); });
    auto fvl = col([](tCNode *c) { return c->getFlowEdg()->getVEdgLen(); });
    auto fdest = icol([&](tCNode *c) { return idx(c->getFlowEdg()->getDestinationPtrNC()); });
    auto snode = icol([&](tCNode *c) { return c->getStreamNode() ? idx((tNode *)c->getStreamNode()) : -1; });
    auto hillpath = col([](tCNode *c) { return c->getHillPath(); });
    auto streampath = col([](tCNode *c) { return c->getStreamPath(); });
    auto carea = col([](tCNode *c) { return c->getContrArea(); });
    auto ttime = col([](tCNode *c) { return c->getTTime(); });
    auto reach = icol([](tCNode *c) { return c->getReach(); });
    std::unordered_map<tEdge *, int> epos;
    { tMeshListIter<tEdge> eit(mesh->getEdgeList()); int k = 0; for (tEdge *e = eit.FirstP(); eit.IsActive(); e = eit.NextP()) epos[e] = k++; }
    std::vector<int32_t> rowptr(n + 1, 0), ncol, lpos;
    std::vector<double> elen, evl, clen, cvl;
    for (int i = 0; i < n; i++) {
      tEdge *first = cells[i]->getEdg(), *e = first;
      do {
        ncol.push_back(idx(e->getDestinationPtrNC()));
        elen.push_back(e->getLength()); evl.push_back(e->getVEdgLen());
        auto ep = epos.find(e); lpos.push_back(ep == epos.end() ? -1 : ep->second);
        tEdge *c = e->FindComplement();
        clen.push_back(c ? c->getLength() : e->getLength()); cvl.push_back(c ? c->getVEdgLen() : e->getVEdgLen());
        e = e->getCCWEdg();
      } while (e != first);
      rowptr[i + 1] = (int32_t)ncol.size();
    }
    tribs_mesh_t M{};
    M.n_cells = n; M.n_edges = (int32_t)ncol.size();
    M.boundary = boundary.data(); M.z = z.data(); M.varea = varea.data(); M.flow_slope = fslope.data();
    M.flow_vedglen = fvl.data(); M.flow_dest = fdest.data(); M.stream_node = snode.data(); M.hillpath = hillpath.data();
    M.streampath = streampath.data(); M.contr_area = carea.data(); M.ttime = ttime.data(); M.reach = reach.data();
    M.nbr_rowptr = rowptr.data(); M.nbr_col = ncol.data(); M.nbr_len = elen.data(); M.nbr_vedglen = evl.data();
    M.nbr_len_in = clen.data(); M.nbr_vedglen_in = cvl.data(); M.nbr_listpos = lpos.data();

    // --- parameters (members of tHydroModel / tFlowNet / tKinemat / tRunTimer)
    auto Ks = col([](tCNode *c) { return c->getKs(); }), ThS = col([](tCNode *c) { return c->getThetaS(); });
    auto ThR = col([](tCNode *c) { return c->getThetaR(); }), Lam = col([](tCNode *c) { return c->getPoreSize(); });
    auto Psib = col([](tCNode *c) { return c->getAirEBubPres(); }), F = col([](tCNode *c) { return c->getDecayF(); });
    auto Ar = col([](tCNode *c) { return c->getSatAnRatio(); }), UAr = col([](tCNode *c) { return c->getUnsatAnRatio(); });
    auto bedrock = col([](tCNode *c) { return c->getBedrockDepth(); });
    tribs_params_t P{};
    P.Ks = Ks.data(); P.ThetaS = ThS.data(); P.ThetaR = ThR.data(); P.PoreSize = Lam.data(); P.AirEBubPres = Psib.data();
    P.DecayF = F.data(); P.SatAnRatio = Ar.data(); P.UnsatAnRatio = UAr.data(); P.bedrock = bedrock.data();
    P.BasArea = hydro->BasArea; P.BasArea_flow = flow->BasArea; P.IntStormMAX = hydro->IntStormMAX;
    P.surfaceSoilDepth = hydro->surfaceSoilDepth; P.rootZoneDepth = hydro->rootZoneDepth;
    P.EToption = hydro->EToption; P.Ioption = hydro->Ioption; P.SnOpt = hydro->SnOpt; P.RdstrOption = hydro->RdstrOption;
    P.velcoef = flow->velcoef; P.velratio = flow->velratio; P.flowexp = flow->flowexp; P.baseflow = flow->baseflow;
    P.kincoef = flow->kincoef; P.dtReff = flow->dtReff; P.outlet_contr_area = flow->OutletNode->getContrArea();
    P.tstep = timer->getTimeStep(); P.gwstep = timer->getGWTimeStep(); P.etistep = timer->getEtIStep();
    P.output_interval = timer->getOutputInterval(); P.res_limit = flow->res->limit;

    // --- initial state: whatever InitSet left in the nodes
    stage.assign(TRIBS_N_FIELDS, std::vector<double>());
    tribs_state_t S{};
    auto put = [&](int f, std::vector<double> v) { stage[f] = std::move(v); S.f[f] = stage[f].data(); };
    put(TRIBS_F_NwtOld, col([](tCNode *c) { return c->getNwtOld(); })); put(TRIBS_F_NwtNew, col([](tCNode *c) { return c->getNwtNew(); }));
    put(TRIBS_F_MuOld, col([](tCNode *c) { return c->getMuOld(); }));   put(TRIBS_F_MuNew, col([](tCNode *c) { return c->getMuNew(); }));
    put(TRIBS_F_MiOld, col([](tCNode *c) { return c->getMiOld(); }));   put(TRIBS_F_MiNew, col([](tCNode *c) { return c->getMiNew(); }));
    put(TRIBS_F_NtOld, col([](tCNode *c) { return c->getNtOld(); }));   put(TRIBS_F_NtNew, col([](tCNode *c) { return c->getNtNew(); }));
    put(TRIBS_F_NfOld, col([](tCNode *c) { return c->getNfOld(); }));   put(TRIBS_F_NfNew, col([](tCNode *c) { return c->getNfNew(); }));
    put(TRIBS_F_RuOld, col([](tCNode *c) { return c->getRuOld(); }));   put(TRIBS_F_RuNew, col([](tCNode *c) { return c->getRuNew(); }));
    put(TRIBS_F_RiOld, col([](tCNode *c) { return c->getRiOld(); }));   put(TRIBS_F_RiNew, col([](tCNode *c) { return c->getRiNew(); }));
    put(TRIBS_F_SoilMoisture, col([](tCNode *c) { return c->getSoilMoisture(); }));
    put(TRIBS_F_SoilMoistureSC, col([](tCNode *c) { return c->getSoilMoistureSC(); }));
    put(TRIBS_F_SoilMoistureUNSC, col([](tCNode *c) { return c->getSoilMoistureUNSC(); }));
    put(TRIBS_F_RootMoisture, col([](tCNode *c) { return c->getRootMoisture(); }));
    put(TRIBS_F_RootMoistureSC, col([](tCNode *c) { return c->getRootMoistureSC(); }));
    put(TRIBS_F_AvSoilMoisture, col([](tCNode *c) { return c->getAvSoilMoisture(); }));
    put(TRIBS_F_TTime, ttime);
    if (nranks <= 1) {
      if (p_init(&M, &P, &S, nullptr, &h)) die("tribs_gpu_init");
      rank_h.assign(1, h);
      owner.assign(n, 0);
    } else {
      p_run = (decltype(p_run))dlsym(lib, "tribs_gpu_run");
      p_wait = (decltype(p_wait))dlsym(lib, "tribs_gpu_wait");
      p_attach_local = (decltype(p_attach_local))dlsym(lib, "tribs_gpu_attach_local");
      p_ipc_export = (decltype(p_ipc_export))dlsym(lib, "tribs_gpu_ipc_export");
      p_attach_peers = (decltype(p_attach_peers))dlsym(lib, "tribs_gpu_attach_peers");
      if (!p_run || !p_wait || !p_attach_local || !p_ipc_export || !p_attach_peers || !cell_owner) die("dlsym (partitions)");
      owner = *cell_owner;
      for (int r = 0; r < nranks; r++) {
        tribs_part_t pt{};
        pt.rank = r; pt.nranks = nranks; pt.n_parts = nranks; pt.max_batch_steps = max_batch_steps; pt.cell_owner = owner.data();
        tribs_handle_t hr = nullptr;
        if (p_init(&M, &P, &S, &pt, &hr)) die("tribs_gpu_init (partition)");
        rank_h.push_back(hr);
      }
      if (p_attach_local(rank_h.data(), nranks)) die("tribs_gpu_attach_local");
      h = rank_h[0];
    }
    for (auto &v : stage) v.assign(n, 0.0);
    rain.assign(n, 0.0);
  }

  // k whole time steps on every rank of this process in one pipelined batch (tribs_gpu_run): the calls return
  // at once, the ranks meet on the device; the host waits once at the end. steps[i] as flags() would fill it
  // for step i (the host advances tRunTimer and tRainfall for the k steps first).
  void runBatch(const std::vector<tribs_step_t> &steps, const std::vector<uint32_t> &masks) {
    if (!p_run) { p_run = (decltype(p_run))dlsym(lib, "tribs_gpu_run"); p_wait = (decltype(p_wait))dlsym(lib, "tribs_gpu_wait"); }
    if (!p_run || !p_wait) die("dlsym (tribs_gpu_run)");
    for (tribs_handle_t hr : rank_h)
      if (p_run(hr, (int)steps.size(), steps.data(), masks.data())) die("tribs_gpu_run");
    for (tribs_handle_t hr : rank_h)
      if (p_wait(hr)) die("tribs_gpu_wait");
  }

  tribs_step_t flags() {
    tribs_step_t s{};
    s.current_time = timer->getCurrentTime();
    s.dt = timer->getTimeStep();
    s.dt_gw = timer->getGWTimeStep();
    s.elapsed_steps = timer->getElapsedSteps(timer->getCurrentTime());
    s.hourly_reset = !(fmod((timer->getCurrentTime() - timer->getTimeStep()), timer->getEtIStep()));   // tHydroModel.cpp:700
    return s;
  }
  // tHydroModel::UnSaturatedZone(double dt)
  void UnSaturatedZone(double) {
    tribs_step_t s = flags();
    for (size_t i = 0; i < cells.size(); i++) rain[i] = cells[i]->getRain();   // whatever tRainfall/tStorm wrote
    s.rain_mode = TRIBS_RAIN_PER_CELL; s.rain = rain.data();
    if (p_step(h, TRIBS_PHASE_UNSAT, &s)) die("UnSaturatedZone");
  }
  void ResetGW() {}   // folded into the SAT phase (a pointer swap on the device)
  // tHydroModel::SaturatedZone(double dtGW)
  void SaturatedZone(double) { tribs_step_t s = flags(); if (p_step(h, TRIBS_PHASE_SAT, &s)) die("SaturatedZone"); }
  // tHydroModel::Reset()
  void Reset() { tribs_step_t s = flags(); if (p_step(h, TRIBS_PHASE_RESET, &s)) die("Reset"); }

  // the fields tKinemat / tFlowNet / tCOutput read each step
  void syncToNodes() {
    const int fields[] = {TRIBS_F_NwtNew, TRIBS_F_MuNew, TRIBS_F_MiNew, TRIBS_F_NtNew, TRIBS_F_NfNew, TRIBS_F_RuNew, TRIBS_F_RiNew,
                          TRIBS_F_Qpin, TRIBS_F_Qpout, TRIBS_F_intstorm, TRIBS_F_srf, TRIBS_F_hsrf, TRIBS_F_sbsrf, TRIBS_F_psrf,
                          TRIBS_F_satsrf, TRIBS_F_esrf, TRIBS_F_srf_hr, TRIBS_F_SoilMoisture, TRIBS_F_SoilMoistureSC,
                          TRIBS_F_SoilMoistureUNSC, TRIBS_F_RootMoisture, TRIBS_F_RootMoistureSC, TRIBS_F_Recharge};
    tribs_state_t S{};
    uint64_t mask = 0;
    for (int f : fields) { S.f[f] = stage[f].data(); mask |= TRIBS_MASK(f); }
   for (size_t r = 0; r < rank_h.size(); r++) {       // every rank of this process hands back the cells it owns
    if (p_sync(rank_h[r], mask, &S)) die("tribs_gpu_sync");
    for (size_t i = 0; i < cells.size(); i++) {
      if (owner[i] != (int32_t)r) continue;
      tCNode *c = cells[i];
      c->setNwtNew(stage[TRIBS_F_NwtNew][i]); c->setMuNew(stage[TRIBS_F_MuNew][i]); c->setMiNew(stage[TRIBS_F_MiNew][i]);
      c->setNtNew(stage[TRIBS_F_NtNew][i]); c->setNfNew(stage[TRIBS_F_NfNew][i]); c->setRuNew(stage[TRIBS_F_RuNew][i]);
      c->setRiNew(stage[TRIBS_F_RiNew][i]); c->setQpin(stage[TRIBS_F_Qpin][i]); c->setQpout(stage[TRIBS_F_Qpout][i]);
      c->setIntStormVar(stage[TRIBS_F_intstorm][i]); c->setsrf(stage[TRIBS_F_srf][i]); c->sethsrf(stage[TRIBS_F_hsrf][i]);
      c->setsbsrf(stage[TRIBS_F_sbsrf][i]); c->setpsrf(stage[TRIBS_F_psrf][i]); c->setsatsrf(stage[TRIBS_F_satsrf][i]);
      c->setesrf(stage[TRIBS_F_esrf][i]); c->setSrf_Hr(stage[TRIBS_F_srf_hr][i]);
      c->setSoilMoisture(stage[TRIBS_F_SoilMoisture][i]); c->setSoilMoistureSC(stage[TRIBS_F_SoilMoistureSC][i]);
      c->setSoilMoistureUNSC(stage[TRIBS_F_SoilMoistureUNSC][i]); c->setRootMoisture(stage[TRIBS_F_RootMoisture][i]);
      c->setRootMoistureSC(stage[TRIBS_F_RootMoistureSC][i]); c->setRecharge(stage[TRIBS_F_Recharge][i]);
    }
   }
  }
  // ------------------------------------------------------------------ channel routing
  // attachChannels hands tKinemat's reach lists, channel geometry and levels to the device (what InitializeStreamReach,
  // tKinemat.cpp:1136-1252, reads per reach and step); routeChannels(k) replaces the reach loop of
  // tKinemat::RunRoutingModel (823-891) for the k steps of the last batch (k = 0: the single step whose ROUTE phase
  // ran last) and writes Hlevel / Qstrm / FlowVelocity back to the stream nodes, OutletHlev and Qout to tKinemat.
  // Inside the reference this is member code of tKinemat (it uses its protected lists).
  tKinemat *kin = nullptr;
  decltype(&tribs_gpu_channel_init) p_ch_init = nullptr;
  decltype(&tribs_gpu_channel_route) p_ch_route = nullptr;
  decltype(&tribs_gpu_channel_sync) p_ch_sync = nullptr;
  std::vector<int32_t> ch_off, ch_node, ch_stream_cells;
  std::vector<double> ch_width, ch_length, ch_slope;

  void attachChannels(tKinemat *flow) {
    kin = flow;
    p_ch_init = (decltype(p_ch_init))dlsym(lib, "tribs_gpu_channel_init");
    p_ch_route = (decltype(p_ch_route))dlsym(lib, "tribs_gpu_channel_route");
    p_ch_sync = (decltype(p_ch_sync))dlsym(lib, "tribs_gpu_channel_sync");
    auto p_n_stream = (decltype(&tribs_gpu_n_stream))dlsym(lib, "tribs_gpu_n_stream");
    auto p_stream_cells = (decltype(&tribs_gpu_stream_cells))dlsym(lib, "tribs_gpu_stream_cells");
    if (!p_ch_init || !p_ch_route || !p_ch_sync || !p_n_stream || !p_stream_cells) die("dlsym (tribs_gpu_channel_*)");
    const int n = (int)cells.size();
    tPtrListIter<tCNode> hi(flow->NodesLstH);
    tListIter<int> ni(flow->NNodes);
    ch_off.assign(1, 0);
    ni.First();
    for (tCNode *head = hi.FirstP(); !hi.AtEnd(); head = hi.NextP(), ni.Next()) {
      tCNode *c = head;
      const int nn = ni.DatRef();
      for (int k = 0; k < nn; k++) {
        const int id = idx((tNode *)c);                       // the basin outlet is not an active cell
        const bool last = k == nn - 1;
        ch_node.push_back(id >= 0 ? id : n);
        ch_width.push_back(c->getChannelWidth());
        ch_length.push_back(last ? 0.0 : c->getFlowEdg()->getLength());
        ch_slope.push_back(last ? 0.0 : c->getFlowEdg()->getSlope());
        if (!last) c = c->getDownstrmNbr();
      }
      ch_off.push_back((int32_t)ch_node.size());
    }
    std::vector<double> h0(n + 1), q0(n + 1), oh(ch_off.size() - 1);
    for (int i = 0; i < n; i++) { h0[i] = cells[i]->getHlevel(); q0[i] = cells[i]->getQstrm(); }
    h0[n] = flow->OutletNode->getHlevel(); q0[n] = flow->OutletNode->getQstrm();
    for (size_t r = 0; r < oh.size(); r++) oh[r] = flow->OutletHlev[r];
    tribs_channel_t C{};
    C.n_reach = (int32_t)oh.size();
    C.percolation_option = flow->percolationOption; C.reservoir_option = flow->optres;
    C.reach_off = ch_off.data(); C.reach_node = ch_node.data();
    C.width = ch_width.data(); C.length = ch_length.data(); C.slope = ch_slope.data();
    C.roughness = flow->Roughness; C.uniform_width = flow->Width; C.dt_s = timer->getTimeStep() * 3600.;
    C.hlevel0 = h0.data(); C.qstrm0 = q0.data(); C.outlet_hlev0 = oh.data();
    for (tribs_handle_t hr : rank_h)
      if (p_ch_init(hr, &C)) die("tribs_gpu_channel_init");
    ch_stream_cells.resize(p_n_stream(rank_h[0]));
    if (p_stream_cells(rank_h[0], ch_stream_cells.data())) die("tribs_gpu_stream_cells");
  }

  // returns tKinemat::Qout after each of the k steps
  std::vector<double> routeChannels(int k) {
    std::vector<double> q((size_t)std::max(k, 1));
    for (tribs_handle_t hr : rank_h)            // every rank routes the whole net from the combined inflows: same result
      if (p_ch_route(hr, k, q.data())) die("tribs_gpu_channel_route");
    const size_t S = ch_stream_cells.size();
    std::vector<double> hl(S + 1), qs(S + 1), vel(S + 1), oh(ch_off.size() - 1);
    if (p_ch_sync(rank_h[0], hl.data(), qs.data(), vel.data(), oh.data(), nullptr)) die("tribs_gpu_channel_sync");
    for (size_t s = 0; s < S; s++) {
      tCNode *c = cells[ch_stream_cells[s]];
      c->setHlevel(hl[s]); c->setQstrm(qs[s]); c->setFlowVelocity(vel[s]);
    }
    kin->OutletNode->setHlevel(hl[S]); kin->OutletNode->setQstrm(qs[S]);
    for (size_t r = 0; r < oh.size(); r++) kin->OutletHlev[r] = oh[r];
    kin->Qout = q.back();
    return q;
  }

  // ------------------------------------------------------------------ kernel (1): ET + interception
  // These bodies replace the cell loops of tEvapoTrans::callEvapoPotential (tEvapoTrans.cpp:603-795)
  // and tEvapoTrans::callEvapoTrans (1010-1028); inside the reference they are member code of
  // tEvapoTrans (they use its protected members).
  tEvapoTrans *et = nullptr;
  tIntercept *icp = nullptr;
  decltype(&tribs_gpu_et_init) p_et_init = nullptr;
  decltype(&tribs_gpu_et_step) p_et_step = nullptr;
  decltype(&tribs_gpu_et_sync) p_et_sync = nullptr;
  std::vector<int32_t> cell_record;
  std::vector<double> met_a[9], met_b[9];
  std::vector<std::vector<double>> et_stage;
  tribs_met_t met_pot{}, met_cmp{};
  tribs_et_step_t es{};

  void attachET(tEvapoTrans *e, tIntercept *ic, tHydroModel *hydro) {
    et = e; icp = ic;
    p_et_init = (decltype(p_et_init))dlsym(lib, "tribs_gpu_et_init");
    p_et_step = (decltype(p_et_step))dlsym(lib, "tribs_gpu_et_step");
    p_et_sync = (decltype(p_et_sync))dlsym(lib, "tribs_gpu_et_sync");
    if (!p_et_init || !p_et_step || !p_et_sync) die("dlsym (et)");
    const int n = (int)cells.size();
    auto col = [&](auto fn) { std::vector<double> v(n); for (int i = 0; i < n; i++) v[i] = fn(cells[i]); return v; };
    auto fslope = col([](tCNode *c) { return c->getFlowEdg()->getSlope(); });
    auto aspect = col([](tCNode *c) { return c->getAspect(); });
    auto hcond = col([](tCNode *c) { return c->getVolHeatCond(); });
    auto hcap = col([](tCNode *c) { return c->getSoilHeatCap(); });
    auto scut = col([](tCNode *c) { return c->getSoilCutoff(); });
    auto rcut = col([](tCNode *c) { return c->getRootCutoff(); });
    std::vector<int32_t> lu(n);
    int maxc = 0;
    for (int i = 0; i < n; i++) { lu[i] = cells[i]->getLandUse(); maxc = std::max(maxc, (int)lu[i]); }
    std::vector<double> table((size_t)(maxc + 1) * 15, 0.0);
    for (int i = 0; i < n; i++) {                                  // tInvariant rows of the classes in use
      hydro->landPtr->setLandPtr(lu[i]);
      for (int k = 1; k <= 14; k++) table[(size_t)lu[i] * 15 + k] = hydro->landPtr->getLandProp(k);
    }
    tribs_et_params_t P{};
    P.flow_slope = fslope.data(); P.aspect = aspect.data(); P.vol_heat_cond = hcond.data(); P.soil_heat_cap = hcap.data();
    P.soil_cutoff = scut.data(); P.root_cutoff = rcut.data(); P.land_class = lu.data(); P.n_land_classes = maxc + 1;
    P.land_table = table.data();
    std::vector<double> shelt, hor0;                       // OPTRADSHELT 1-3: what tShelter left in the nodes
    std::vector<const double *> init0(TRIBS_N_ET_FIELDS, nullptr);
    if (et->shelterOption >= 1 && et->shelterOption <= 3) {
      shelt = col([](tCNode *c) { return c->getSheltFact(); });
      hor0 = col([](tCNode *c) { return c->getHorAngle0000(); });
      init0[TRIBS_ET_SheltFact] = shelt.data();
      P.init = init0.data(); P.hor_angle_0000 = hor0.data();
    }
    P.et_option = et->evapotransOption; P.i_option = et->Ioption; P.gflux_option = et->gFluxOption;
    P.shelter_option = et->shelterOption; P.lu_option = et->luOption; P.snow_option = et->snowOption;
    P.metstep_min = et->timeStep; P.etistep_h = timer->getEtIStep(); P.tlinke = et->Tlinke;
    P.temp_lapse = et->tempLapseRate; P.max_interstorm = icp->maxInterStormPeriod;
    if (p_et_init(h, &P)) die("tribs_gpu_et_init");
    if (et->metdataOption != 1) die("only METDATAOPTION 1 is wired in this binding");
    cell_record.resize(n);
    for (int i = 0; i < n; i++)
      for (int k = 0; k < et->numStations; k++)
        if (et->assignedStation[i] == et->weatherStations[k].getStation()) cell_record[i] = k;
    et_stage.assign(TRIBS_N_ET_FIELDS, std::vector<double>(n, 0.0));
  }

  void fillMet(int t, std::vector<double> *v, tribs_met_t &m) {      // newHydroMetData(t) for every station
    const int ns = et->numStations;
    for (int q = 0; q < 9; q++) v[q].resize(ns);
    for (int k = 0; k < ns; k++) {
      tHydroMet &st = et->weatherStations[k];
      v[0][k] = st.getAirTemp(t); v[1][k] = st.getDewTemp(t); v[2][k] = st.getRHumidity(t); v[3][k] = st.getVaporPress(t);
      v[4][k] = st.getAtmPress(t); v[5][k] = st.getWindSpeed(t); v[6][k] = st.getSkyCover(t); v[7][k] = st.getRadGlobal(t);
      v[8][k] = st.getOther();
    }
    m.n_records = ns; m.cell_record = cell_record.data();
    m.air_temp = v[0].data(); m.dew_temp = v[1].data(); m.rel_humid = v[2].data(); m.vap_press = v[3].data();
    m.atm_press = v[4].data(); m.wind_speed = v[5].data(); m.sky_cover = v[6].data(); m.rad_global = v[7].data();
    m.elev = v[8].data();
  }

  void sunAndClock() {
    es.alphaD = et->alphaD; es.sinAlpha = et->sinAlpha; es.sunaz = et->sunaz; es.Io = et->Io;
    es.hour = et->currentTime[3]; es.dew_hum_flag = et->dewHumFlag; es.ts_option = et->tsOption;
  }

  // tEvapoTrans::callEvapoPotential() (also the head and tail of tSnowPack::callSnowPack, tSnowPack.cpp:249-285, 731-733)
  void callEvapoPotential() {
    et->SetEnvironment();                                  // calendar + sun geometry, once per call
    const int n = (int)cells.size(), t = et->hourlyTimeStep;
    tribs_step_t s = flags();                              // the sweep reads this step's rain
    for (int i = 0; i < n; i++) rain[i] = cells[i]->getRain();
    s.rain_mode = TRIBS_RAIN_PER_CELL; s.rain = rain.data();
    if (p_step(h, 0, &s)) die("rain upload");
    fillMet(t, met_a, met_pot);
    if (t == 0) {                                          // record-0 decisions, tEvapoTrans.cpp:3989-4003
      tHydroMet &st = et->weatherStations[cell_record[n - 1]];
      et->latitude = st.getLat(1); et->longitude = st.getLong(1); et->gmt = st.getGmt();
      const int k = cell_record[n - 1];
      auto nd = [](double v) { return fabs(v - 9999.99) < 1.0E-3; };
      if (nd(met_a[1][k]) && nd(met_a[3][k])) et->dewHumFlag = 0;
      else if (nd(met_a[2][k]) && nd(met_a[3][k])) et->dewHumFlag = 1;
      else if (nd(met_a[2][k]) && nd(met_a[1][k])) et->dewHumFlag = 2;
    }
    int sky_from = et->skycover_flag ? 0 : n;              // sticky no-data switch, tEvapoTrans.cpp:724-727
    if (!et->skycover_flag)
      for (int i = 0; i < n; i++)
        if (fabs(met_a[6][cell_record[i]] - 9999.99) < 1.0E-3) { sky_from = i; et->skycover_flag = 1; break; }
    sunAndClock();
    es.do_potential = 1; es.first_call = (t == 0); es.sky_flag_from = sky_from;
    es.te_met = (double)timer->getElapsedMETSteps(timer->getCurrentTime());
    es.met_potential = &met_pot;
    et->timeCount++; et->oldTimeStep = et->hourlyTimeStep; et->hourlyTimeStep++;
  }
  // tEvapoTrans::callEvapoTrans(tIntercept*, int flag)
  void callEvapoTrans(int flag) {
    fillMet(et->hourlyTimeStep, met_b, met_cmp);
    sunAndClock();
    es.do_components = 1; es.intercept_flag = flag;
    es.te_eti = (double)timer->getElapsedETISteps(timer->getCurrentTime());
    es.met_components = &met_cmp;
    if (!es.do_potential) es.sky_flag_from = et->skycover_flag ? 0 : (int)cells.size();
    et->timeCount++;
    flushET();
  }
  // one fused launch for the sweeps queued in this step (Simulator::SurfaceHydroProcesses)
  void flushET() {
    if (!es.do_potential && !es.do_components) return;
    if (p_et_step(h, &es)) die("tribs_gpu_et_step");
    es = tribs_et_step_t{};
  }
  // what tCOutput / tWaterBalance read from the nodes (only needed when an output is due)
  void syncEtToNodes() {
    std::vector<double *> ptr(TRIBS_N_ET_FIELDS);
    for (int f = 0; f < TRIBS_N_ET_FIELDS; f++) ptr[f] = et_stage[f].data();
    if (p_et_sync(h, ~0ull, ptr.data())) die("tribs_gpu_et_sync");
    const int mf[] = {TRIBS_F_EvapSoil, TRIBS_F_EvapDryCanopy, TRIBS_F_NetPrecipitation};
    tribs_state_t S{};
    uint64_t mask = 0;
    for (int f : mf) { S.f[f] = stage[f].data(); mask |= TRIBS_MASK(f); }
    if (p_sync(h, mask, &S)) die("tribs_gpu_sync");
    auto &E = et_stage;
    for (size_t i = 0; i < cells.size(); i++) {
      tCNode *c = cells[i];
      c->setPotEvap(E[TRIBS_ET_PotEvap][i]); c->setActEvap(E[TRIBS_ET_ActEvap][i]);
      c->setSurfTemp(E[TRIBS_ET_SurfTemp][i]); c->setSoilTemp(E[TRIBS_ET_SoilTemp][i]);
      c->setNetRad(E[TRIBS_ET_NetRad][i]); c->setGFlux(E[TRIBS_ET_GFlux][i]); c->setHFlux(E[TRIBS_ET_HFlux][i]);
      c->setLFlux(E[TRIBS_ET_LFlux][i]); c->setLongRadIn(E[TRIBS_ET_LongRadIn][i]); c->setLongRadOut(E[TRIBS_ET_LongRadOut][i]);
      c->setShortRadIn(E[TRIBS_ET_ShortRadIn][i]); c->setShortRadIn_dir(E[TRIBS_ET_ShortRadIn_dir][i]);
      c->setShortRadIn_dif(E[TRIBS_ET_ShortRadIn_dif][i]); c->setShortRadSlope(E[TRIBS_ET_ShortRadSlope][i]);
      c->setShortAbsbVeg(E[TRIBS_ET_ShortAbsbVeg][i]); c->setShortAbsbSoi(E[TRIBS_ET_ShortAbsbSoi][i]);
      c->setAirTemp(E[TRIBS_ET_AirTemp][i]); c->setDewTemp(E[TRIBS_ET_DewTemp][i]); c->setRelHumid(E[TRIBS_ET_RelHumid][i]);
      c->setVapPressure(E[TRIBS_ET_VapPressure][i]); c->setSkyCover(E[TRIBS_ET_SkyCover][i]);
      c->setWindSpeed(E[TRIBS_ET_WindSpeed][i]); c->setAirPressure(E[TRIBS_ET_AirPressure][i]);
      c->addCumHrsSun(E[TRIBS_ET_CumHrsSun][i] - c->getCumHrsSun());
      c->setAvEvapFract(E[TRIBS_ET_AvEvapFract][i]); c->setSheltFact(E[TRIBS_ET_SheltFact][i]);
      c->addRSin(E[TRIBS_ET_CumRSin][i] - c->getCumRSin());
      c->setEvapWetCanopy(E[TRIBS_ET_EvapWetCanopy][i]); c->setEvapoTrans(E[TRIBS_ET_EvapoTrans][i]);
      c->addTotEvap(E[TRIBS_ET_CumTotEvap][i] - c->getCumTotEvap());
      c->addBarEvap(E[TRIBS_ET_CumBarEvap][i] - c->getCumBarEvap());
      c->setAvET(E[TRIBS_ET_AvET][i]); c->setInterceptLoss(E[TRIBS_ET_InterceptLoss][i]);
      c->setCanStorage(E[TRIBS_ET_CanStorage][i]); c->setCumIntercept(E[TRIBS_ET_CumIntercept][i]);
      c->setStormLength(0, 0.0); c->setStormLength(1, E[TRIBS_ET_StormLength][i]);
      c->setEvapSoil(stage[TRIBS_F_EvapSoil][i]); c->setEvapDryCanopy(stage[TRIBS_F_EvapDryCanopy][i]);
      c->setNetPrecipitation(stage[TRIBS_F_NetPrecipitation][i]);
    }
  }
  // ------------------------------------------------------------------ snow pack (OPTSNOW 1)
  // tSnowPack derives from tEvapoTrans; under OPTSNOW the simulator calls SnowPack->callSnowPack(Intercept, flag)
  // at every met hour instead of the two ET sweeps (tSimul.cpp:464-492). attachSnow() = attachET() on that object
  // + tribs_gpu_snow_init; callSnowPack() = the call's host part (SetEnvironment, the hour's station records,
  // the record-0 decisions, the clocks) + one tribs_gpu_snow_step for the cell loop (tSnowPack.cpp:286-725).
  tSnowPack *snow = nullptr;
  decltype(&tribs_gpu_snow_init) p_snow_init = nullptr;
  decltype(&tribs_gpu_snow_step) p_snow_step = nullptr;
  decltype(&tribs_gpu_snow_sync) p_snow_sync = nullptr;
  std::vector<std::vector<double>> snow_stage;

  void attachSnow(tSnowPack *sp, tIntercept *ic, tHydroModel *hydro) {
    snow = sp;
    attachET(sp, ic, hydro);                               // the tEvapoTrans members of the tSnowPack object
    p_snow_init = (decltype(p_snow_init))dlsym(lib, "tribs_gpu_snow_init");
    p_snow_step = (decltype(p_snow_step))dlsym(lib, "tribs_gpu_snow_step");
    p_snow_sync = (decltype(p_snow_sync))dlsym(lib, "tribs_gpu_snow_sync");
    if (!p_snow_init || !p_snow_step || !p_snow_sync) die("dlsym (snow)");
    tribs_snow_params_t P{};
    P.min_sn_temp = sp->minSnTemp; P.snliqfrac = sp->snliqfrac; P.hill_albedo_option = sp->hillAlbedoOption;
    if (p_snow_init(h, &P)) die("tribs_gpu_snow_init");
    snow_stage.assign(TRIBS_N_SNOW_FIELDS, std::vector<double>(cells.size(), 0.0));
  }
  // tSnowPack::callSnowPack(tIntercept*, int flag); flag = (Ioption != 0), which the library knows from et_init
  void callSnowPack(int) {
    const int t = et->hourlyTimeStep;                      // tEvapoTrans::hourlyTimeStep of this call
    callEvapoPotential();                                  // fills `es` (do_potential, met record, clocks) and advances them
    es.te_eti = (double)timer->getElapsedETISteps(timer->getCurrentTime());
    if (p_snow_step(h, &es, t)) die("tribs_gpu_snow_step");
    es = tribs_et_step_t{};
  }
  void syncSnowToNodes() {
    syncEtToNodes();
    std::vector<double *> ptr(TRIBS_N_SNOW_FIELDS);
    for (int f = 0; f < TRIBS_N_SNOW_FIELDS; f++) ptr[f] = snow_stage[f].data();
    if (p_snow_sync(h, ~0ull, ptr.data())) die("tribs_gpu_snow_sync");
    const int mf[] = {TRIBS_F_LiqWE, TRIBS_F_IceWE, TRIBS_F_LiqRouted};
    tribs_state_t S{};
    uint64_t mask = 0;
    for (int f : mf) { S.f[f] = stage[f].data(); mask |= TRIBS_MASK(f); }
    if (p_sync(h, mask, &S)) die("tribs_gpu_sync");
    auto &W = snow_stage;
    for (size_t i = 0; i < cells.size(); i++) {
      tCNode *c = cells[i];
      c->setLiqWE(stage[TRIBS_F_LiqWE][i]); c->setIceWE(stage[TRIBS_F_IceWE][i]); c->setLiqRouted(stage[TRIBS_F_LiqRouted][i]);
      c->setSnTempC(W[TRIBS_SN_SnTempC][i]); c->setCrustAge(W[TRIBS_SN_CrustAge][i]); c->setDensityAge(W[TRIBS_SN_DensityAge][i]);
      c->setEvapoTransAge(W[TRIBS_SN_EvapoTransAge][i]); c->setSnSub(W[TRIBS_SN_SnSub][i]); c->setSnEvap(W[TRIBS_SN_SnEvap][i]);
      c->setDU(W[TRIBS_SN_DU][i]); c->setUnode(W[TRIBS_SN_Unode][i]); c->setUerror(W[TRIBS_SN_Uerror][i]);
      c->setSnLHF(W[TRIBS_SN_SnLHF][i]); c->setSnSHF(W[TRIBS_SN_SnSHF][i]); c->setSnGHF(W[TRIBS_SN_SnGHF][i]);
      c->setSnPHF(W[TRIBS_SN_SnPHF][i]); c->setSnRLout(W[TRIBS_SN_SnRLout][i]); c->setSnRLin(W[TRIBS_SN_SnRLin][i]);
      c->setSnRSin(W[TRIBS_SN_SnRSin][i]);
      c->setIntSWE(W[TRIBS_SN_IntSWE][i]); c->setIntPrec(W[TRIBS_SN_IntPrec][i]); c->setIntSnUnload(W[TRIBS_SN_IntSnUnload][i]);
      c->setIntSub(W[TRIBS_SN_IntSub][i]);
      c->addIntSub(W[TRIBS_SN_CumIntSub][i] - c->getCumIntSub()); c->addIntUnl(W[TRIBS_SN_CumIntUnl][i] - c->getCumIntUnl());
      c->addLatHF(W[TRIBS_SN_CumLHF][i] - c->getCumLHF()); c->addMelt(W[TRIBS_SN_CumMelt][i] - c->getCumMelt());
      c->addSHF(W[TRIBS_SN_CumSHF][i] - c->getCumSHF()); c->addSnSub(W[TRIBS_SN_CumSnSub][i] - c->getCumSnSub());
      c->addSnEvap(W[TRIBS_SN_CumSnEvap][i] - c->getCumSnEvap()); c->addPHF(W[TRIBS_SN_CumPHF][i] - c->getCumPHF());
      c->addRLin(W[TRIBS_SN_CumRLin][i] - c->getCumRLin()); c->addRLout(W[TRIBS_SN_CumRLout][i] - c->getCumRLout());
      c->addGHF(W[TRIBS_SN_CumGHF][i] - c->getCumGHF()); c->addCumUerror(W[TRIBS_SN_CumUerror][i] - c->getCumUerror());
      c->addCumHrsSnow(W[TRIBS_SN_CumHrsSnow][i] - c->getCumHrsSnow());
      c->setPersTimeMax(W[TRIBS_SN_PersTimeMax][i]); c->setPersTime(W[TRIBS_SN_PersTime][i]); c->setPeakSWE(W[TRIBS_SN_PeakSWE][i]);
      c->setPeakSWETemp(W[TRIBS_SN_PeakSWETemp][i]); c->setInitPackTime(W[TRIBS_SN_InitPackTime][i]);
      c->setInitPackTimeTemp(W[TRIBS_SN_InitPackTimeTemp][i]); c->setPeakPackTime(W[TRIBS_SN_PeakPackTime][i]);
    }
  }
  ~TribsGpuBinding() { if (h && p_destroy) p_destroy(h); }
};
