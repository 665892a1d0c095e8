// checkpoint.cuh — the dynamic state of a handle: an in-memory snapshot (device copies) and a binary
// checkpoint file. (Included by tribs_gpu.cu; not a stand-alone header.)
//
// The reference's restart (tRestart.cpp, tCNode::writeRestart/readRestart) walks the tCNode lists; its
// writer is broken. Here the state is the structure-of-arrays the path owns:
// every dynamic field, the ET / interception and snow fields, the two tSnowPack members that outlive a
// call, the pending dtReff bins of the hillslope routing ((TimeInd, Qeff) stacks), the tFlowResults arrays
// and the basin accumulators. Per-cell arrays are stored in sweep order, so a file can be loaded into a
// handle with another partition layout.
struct StateSnapshot {
  std::vector<double *> cell;          // one n_pad copy per per-cell array, in enumeration order
  double *ring = nullptr, *results = nullptr, *globals = nullptr, *qeff_now = nullptr;
  bool new_valid = true, at_boundary = true, gw_dirty = false;
  double members[2] = {0.88, 0.0};
  bool valid = false;
};

// channel router state (channel.cuh): levels, discharges, velocities per stream slot and OutletHlev per reach
static int channel_snapshot(tribs_ctx *c);
static int channel_rollback(tribs_ctx *c);
static bool channel_present(tribs_ctx *c);
static bool channel_write(tribs_ctx *c, FILE *f);
static int channel_read(tribs_ctx *c, FILE *f);

static void state_arrays(tribs_ctx *c, std::vector<double *> &out) {
  out.clear();
  for (int k = 0; k < TRIBS_N_FIELDS; k++) out.push_back(c->dv.f[k]);
  if (c->et) {
    for (int k = 0; k < TRIBS_N_ET_FIELDS; k++) out.push_back(c->et->ev.f[k]);
    if (c->et->snow_ready) for (int k = 0; k < TRIBS_N_SNOW_FIELDS; k++) out.push_back(c->et->snv.f[k]);
  }
}

extern "C" int tribs_gpu_snapshot(tribs_handle_t c) {
  if (!c) return fail("tribs_gpu_snapshot: null handle");
  CK(cudaSetDevice(c->device));
  if (!c->snap) c->snap = new StateSnapshot();
  StateSnapshot &sn = *c->snap;
  std::vector<double *> arr;
  state_arrays(c, arr);
  const size_t slots = (size_t)c->n_stream + 1;
  if (sn.cell.size() != arr.size()) {
    CK(cudaStreamSynchronize(c->stream));
    sn.cell.resize(arr.size(), nullptr);
    for (auto &p : sn.cell) if (!p && dev_alloc(c, &p, (size_t)c->n_pad)) return 1;
    if (!sn.ring && (dev_alloc(c, &sn.ring, slots * c->rc.ring) || dev_alloc(c, &sn.results, (size_t)TRIBS_N_RESULTS * c->rc.res_limit) ||
                     dev_alloc(c, &sn.globals, (size_t)TRIBS_N_GLOBALS) || dev_alloc(c, &sn.qeff_now, slots)))
      return 1;
  }
  cudaStream_t st = c->stream;
  for (size_t k = 0; k < arr.size(); k++)
    CK(cudaMemcpyAsync(sn.cell[k], arr[k], (size_t)c->n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(sn.ring, c->d_ring, slots * c->rc.ring * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(sn.results, c->d_results, (size_t)TRIBS_N_RESULTS * c->rc.res_limit * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(sn.globals, c->d_globals, TRIBS_N_GLOBALS * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(sn.qeff_now, c->d_qeff_now, slots * sizeof(double), cudaMemcpyDeviceToDevice, st));
  sn.new_valid = c->new_valid; sn.at_boundary = c->at_boundary; sn.gw_dirty = c->gw_dirty;
  if (c->et && c->et->snow_ready && tribs_gpu_snow_members(c, sn.members, nullptr)) return 1;
  if (channel_snapshot(c)) return 1;
  sn.valid = true;
  return 0;
}

extern "C" int tribs_gpu_rollback(tribs_handle_t c) {
  if (!c) return fail("tribs_gpu_rollback: null handle");
  if (!c->snap || !c->snap->valid) return fail("tribs_gpu_rollback: no snapshot taken");
  CK(cudaSetDevice(c->device));
  StateSnapshot &sn = *c->snap;
  std::vector<double *> arr;
  state_arrays(c, arr);
  if (arr.size() != sn.cell.size()) return fail("tribs_gpu_rollback: the handle gained fields since the snapshot");
  const size_t slots = (size_t)c->n_stream + 1;
  cudaStream_t st = c->stream;
  for (size_t k = 0; k < arr.size(); k++)
    CK(cudaMemcpyAsync(arr[k], sn.cell[k], (size_t)c->n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(c->d_ring, sn.ring, slots * c->rc.ring * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(c->d_results, sn.results, (size_t)TRIBS_N_RESULTS * c->rc.res_limit * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(c->d_globals, sn.globals, TRIBS_N_GLOBALS * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(c->d_qeff_now, sn.qeff_now, slots * sizeof(double), cudaMemcpyDeviceToDevice, st));
  c->new_valid = sn.new_valid; c->at_boundary = sn.at_boundary; c->gw_dirty = sn.gw_dirty;
  if (c->et && c->et->snow_ready && tribs_gpu_snow_set_members(c, sn.members)) return 1;
  return channel_rollback(c);
}

// ---- checkpoint file: "TRIBSCK1", int32 {n_cells, n_arrays, n_slots, ring, n_results, res_limit, n_globals, flags},
//      2 doubles (snow members), then n_arrays x n_cells doubles (sweep order), ring, results, globals, qeff_now; with
//      flags bit 2: "CHAN", int32 {n_reach, n_slots}, Hlevel, Qstrm, FlowVelocity per stream slot, OutletHlev per reach
extern "C" int tribs_gpu_save_state(tribs_handle_t c, const char *path) {
  if (!c || !path) return fail("tribs_gpu_save_state: null argument");
  CK(cudaSetDevice(c->device));
  if (check_sweep_abort(c)) return 1;
  if (!c->at_boundary) return fail("tribs_gpu_save_state: call it on a step boundary (after RESET)");
  FILE *f = fopen(path, "wb");
  if (!f) return fail(std::string("tribs_gpu_save_state: cannot open ") + path);
  std::vector<double *> arr;
  state_arrays(c, arr);
  const int32_t slots = c->n_stream + 1;
  const int32_t hdr[8] = {c->n, (int32_t)arr.size(), slots, c->rc.ring, TRIBS_N_RESULTS, c->rc.res_limit, TRIBS_N_GLOBALS,
                          (c->new_valid ? 1 : 0) | (c->gw_dirty ? 2 : 0) | (channel_present(c) ? 4 : 0)};
  double members[2] = {0.88, 0.0};
  if (c->et && c->et->snow_ready && tribs_gpu_snow_members(c, members, nullptr)) { fclose(f); return 1; }
  bool ok = fwrite("TRIBSCK1", 1, 8, f) == 8 && fwrite(hdr, sizeof(int32_t), 8, f) == 8 && fwrite(members, sizeof(double), 2, f) == 2;
  std::vector<double> h((size_t)c->n);
  for (size_t k = 0; ok && k < arr.size(); k++) {
    // Old/New alias after Reset(): what the reference calls New is in the Old buffer then
    const double *src = (k < TRIBS_N_FIELDS) ? field_ptr(c, (int)k) : arr[k];
    gather_to_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(src, c->d_dev_of_sweep, c->d_stage, c->n);
    c->launches++;
    if (cudaMemcpyAsync(h.data(), c->d_stage, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) { ok = false; break; }
    ok = fwrite(h.data(), sizeof(double), h.size(), f) == h.size();
  }
  auto dump = [&](const double *dptr, size_t count) {
    std::vector<double> b(count);
    if (cudaMemcpy(b.data(), dptr, count * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    return fwrite(b.data(), sizeof(double), count, f) == count;
  };
  ok = ok && dump(c->d_ring, (size_t)slots * c->rc.ring) && dump(c->d_results, (size_t)TRIBS_N_RESULTS * c->rc.res_limit) &&
       dump(c->d_globals, TRIBS_N_GLOBALS) && dump(c->d_qeff_now, slots);
  ok = ok && channel_write(c, f);
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : fail(std::string("tribs_gpu_save_state: write to ") + path + " failed");
}

extern "C" int tribs_gpu_load_state(tribs_handle_t c, const char *path) {
  if (!c || !path) return fail("tribs_gpu_load_state: null argument");
  CK(cudaSetDevice(c->device));
  FILE *f = fopen(path, "rb");
  if (!f) return fail(std::string("tribs_gpu_load_state: cannot open ") + path);
  std::vector<double *> arr;
  state_arrays(c, arr);
  const int32_t slots = c->n_stream + 1;
  char magic[8];
  int32_t hdr[8];
  double members[2];
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "TRIBSCK1", 8) == 0 && fread(hdr, sizeof(int32_t), 8, f) == 8 &&
            fread(members, sizeof(double), 2, f) == 2;
  if (ok && (hdr[0] != c->n || hdr[1] != (int32_t)arr.size() || hdr[2] != slots || hdr[3] != c->rc.ring || hdr[4] != TRIBS_N_RESULTS ||
             hdr[5] != c->rc.res_limit || hdr[6] != TRIBS_N_GLOBALS)) {
    fclose(f);
    return fail("tribs_gpu_load_state: the checkpoint belongs to another basin or option set");
  }
  std::vector<double> h((size_t)c->n);
  c->new_valid = true;     // the file holds New and Old separately
  for (size_t k = 0; ok && k < arr.size(); k++) {
    ok = fread(h.data(), sizeof(double), h.size(), f) == h.size();
    if (!ok) break;
    if (cudaMemcpyAsync(c->d_stage, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { ok = false; break; }
    scatter_from_sweep_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->d_stage, c->d_dev_of_sweep, arr[k], c->n);
    c->launches++;
    ok = cudaStreamSynchronize(c->stream) == cudaSuccess;
  }
  auto load = [&](double *dptr, size_t count) {
    std::vector<double> b(count);
    if (fread(b.data(), sizeof(double), count, f) != count) return false;
    return cudaMemcpy(dptr, b.data(), count * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  ok = ok && load(c->d_ring, (size_t)slots * c->rc.ring) && load(c->d_results, (size_t)TRIBS_N_RESULTS * c->rc.res_limit) &&
       load(c->d_globals, TRIBS_N_GLOBALS) && load(c->d_qeff_now, slots);
  if (ok && channel_present(c)) {
    if (!(hdr[7] & 4)) { fclose(f); return fail("tribs_gpu_load_state: the handle routes its channels on the device, the checkpoint holds no channel state"); }
    if (channel_read(c, f)) { fclose(f); return 1; }
  }
  fclose(f);
  if (!ok) return fail(std::string("tribs_gpu_load_state: ") + path + " is truncated or unreadable");
  c->gw_dirty = (hdr[7] & 2) != 0;
  c->at_boundary = true;
  if (c->et && c->et->snow_ready && tribs_gpu_snow_set_members(c, members)) return 1;
  return 0;
}

// pre-size the batch buffers (otherwise the first tribs_gpu_run of a given length allocates them)
extern "C" int tribs_gpu_reserve_batch(tribs_handle_t c, int n_steps) {
  if (!c || n_steps <= 0) return fail("tribs_gpu_reserve_batch: bad argument");
  CK(cudaSetDevice(c->device));
  return ensure_pipe(c, pipe_of(c), n_steps);
}

static void tribs_free_snapshot(tribs_ctx *c) {
  delete c->snap;   // its device buffers live in the handle's allocation list
  c->snap = nullptr;
}
