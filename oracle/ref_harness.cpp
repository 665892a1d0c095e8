// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Reference harness: links the *unmodified* reference sources where they lie under
// /root/reference (see oracle/Makefile; nothing is copied into this repo), wires the same
// objects the reference's serialSimulation() wires (reference: src/main.cpp:64-155), and
// runs the reference's own time loop (reference: src/tSimulator/tSimul.cpp:248-293) one
// call at a time so that raw FP64 state can be dumped between phases. Text outputs of the
// reference carry only 4-7 significant digits, so this is the only way to pin 1e-6
// relative parity.
//
// Outputs (all under --out DIR, container format "TRBDUMP1", see tribs_b200/trb.py):
//   basin.trb  flattened mesh / parameters / initial state in sweep (list) order
//   steps.trb  per-step, per-phase snapshots of the hot-path tCNode fields + global sums
//   kat.trb    known-answer vectors of the reference's scalar helper functions
//   timing: one JSON line on stdout ("ref_timing")
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may execute the binary built from this file.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>
#include <list>
#include <unordered_map>
#include <iostream>
#include <fstream>
#include <sstream>
#include <chrono>
#include <memory>
#include <algorithm>
#include <iomanip>

// The harness needs to read a handful of private accumulators (Stok, TotRain, psrf ...).
// Access specifiers do not change the object layout g++ produces, and the reference
// translation units are compiled untouched.
#define private public
#define protected public
#include "src/Headers/globalIO.h"
#include "Headers/Inclusions.h"
#include "src/Mathutil/predicates.h"
#include "src/tFlowNet/tKinemat.h"
#include "tFlowNet/tReservoir.h"
#include "src/tRasTin/tRainfall.h"
#include "src/tRasTin/tShelter.h"
#include "src/tSimulator/tSimul.h"
#include "src/Headers/TemplDefinitions.h"
#include "src/tHydro/tSnowPack.h"
#include "src/tFlowNet/tFlowResults.h"
#include "src/tGraph/tGraph.h"
// the reference-side binding of the C ABI (what INTEGRATION.md tells a maintainer to add);
// inside tHydroModel it would have member access, here it gets it through the macro above
#include "tribs_ref_binding.h"
#undef private
#undef protected

using namespace std;

Predicates predicate;
tOstream Cout(cout);
tOstream Cerr(cerr);

// ---------------------------------------------------------------------------------
// TRBDUMP1 writer
// ---------------------------------------------------------------------------------
struct TrbWriter {
    FILE *f = nullptr;
    explicit TrbWriter(const string &path) {
        f = fopen(path.c_str(), "wb");
        if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(3); }
        fwrite("TRBDUMP1", 1, 8, f);
    }
    ~TrbWriter() { if (f) fclose(f); }
    void rec(const string &name, int dtype, int64_t n, const void *p, size_t esz) {
        char nm[48]; memset(nm, 0, sizeof nm);
        strncpy(nm, name.c_str(), 47);
        fwrite(nm, 1, 48, f);
        int32_t dt = dtype; fwrite(&dt, 4, 1, f);
        int32_t pad = 0;    fwrite(&pad, 4, 1, f);
        fwrite(&n, 8, 1, f);
        if (n) fwrite(p, esz, (size_t)n, f);
    }
    void f64(const string &name, const vector<double> &v) { rec(name, 0, (int64_t)v.size(), v.data(), 8); }
    void i32(const string &name, const vector<int32_t> &v) { rec(name, 1, (int64_t)v.size(), v.data(), 4); }
    void f64s(const string &name, double v) { rec(name, 0, 1, &v, 8); }
    void i32s(const string &name, int32_t v) { rec(name, 1, 1, &v, 4); }
};

// ---------------------------------------------------------------------------------
struct Harness {
    SimulationControl *graph_sim = nullptr;   // set: dump tGraph's reach partitions with the basin
    tMesh<tCNode> *mesh;
    tKinemat *flow;
    tHydroModel *hydro;
    tRunTimer *timer;
    tEvapoTrans *et = nullptr;
    tIntercept *icp = nullptr;
    tRainfall *rainfall = nullptr;
    tSnowPack *snow = nullptr;
    vector<tCNode*> cells;                    // active cells in list (sweep) order
    unordered_map<tNode*, int> pos;           // node -> index in cells

    void index() {
        tMeshListIter<tCNode> it(mesh->getNodeList());
        tCNode *cn;
        for (cn = it.FirstP(); it.IsActive(); cn = it.NextP()) {
            pos[(tNode*)cn] = (int)cells.size();
            cells.push_back(cn);
        }
    }
    int idx(tNode *n) const {
        auto k = pos.find(n);
        return k == pos.end() ? -1 : k->second;
    }

    template <class F> vector<double> col(F fn) {
        vector<double> v(cells.size());
        for (size_t i = 0; i < cells.size(); i++) v[i] = fn(cells[i]);
        return v;
    }
    template <class F> vector<int32_t> icol(F fn) {
        vector<int32_t> v(cells.size());
        for (size_t i = 0; i < cells.size(); i++) v[i] = fn(cells[i]);
        return v;
    }

    // -------------------------------------------------------------- static basin
    void dump_basin(const string &path, tInputFile &inf) {
        TrbWriter w(path);
        int N = (int)cells.size();
        w.i32s("n_active", N);
        w.i32s("n_nodes_total", mesh->getNodeList()->getSize());
        w.i32("id", icol([](tCNode *c) { return c->getID(); }));
        w.i32("boundary", icol([](tCNode *c) { return c->getBoundaryFlag(); }));
        w.f64("x", col([](tCNode *c) { return c->getX(); }));
        w.f64("y", col([](tCNode *c) { return c->getY(); }));
        w.f64("z", col([](tCNode *c) { return c->getZ(); }));
        w.f64("varea", col([](tCNode *c) { return c->getVArea(); }));
        w.f64("bedrock", col([](tCNode *c) { return c->getBedrockDepth(); }));
        w.f64("flow_slope", col([](tCNode *c) { return c->getFlowEdg()->getSlope(); }));
        w.f64("flow_len", col([](tCNode *c) { return c->getFlowEdg()->getLength(); }));
        w.f64("flow_vedglen", col([](tCNode *c) { return c->getFlowEdg()->getVEdgLen(); }));
        w.i32("flow_dest", icol([this](tCNode *c) { return idx(c->getFlowEdg()->getDestinationPtrNC()); }));
        w.i32("flow_edge_id", icol([](tCNode *c) { return c->getFlowEdg()->getID(); }));
        w.i32("stream_node", icol([this](tCNode *c) { return c->getStreamNode() ? idx((tNode*)c->getStreamNode()) : -1; }));
        w.f64("hillpath", col([](tCNode *c) { return c->getHillPath(); }));
        w.f64("streampath", col([](tCNode *c) { return c->getStreamPath(); }));
        w.f64("ttime", col([](tCNode *c) { return c->getTTime(); }));
        w.f64("contr_area", col([](tCNode *c) { return c->getContrArea(); }));
        w.i32("reach", icol([](tCNode *c) { return c->getReach(); }));
        w.i32("soil_id", icol([](tCNode *c) { return c->getSoilID(); }));
        w.i32("land_use", icol([](tCNode *c) { return c->getLandUse(); }));
        w.f64("aspect", col([](tCNode *c) { return c->getAspect(); }));
        // soil
        w.f64("Ks", col([](tCNode *c) { return c->getKs(); }));
        w.f64("ThetaS", col([](tCNode *c) { return c->getThetaS(); }));
        w.f64("ThetaR", col([](tCNode *c) { return c->getThetaR(); }));
        w.f64("PoreSize", col([](tCNode *c) { return c->getPoreSize(); }));
        w.f64("AirEBubPres", col([](tCNode *c) { return c->getAirEBubPres(); }));
        w.f64("DecayF", col([](tCNode *c) { return c->getDecayF(); }));
        w.f64("SatAnRatio", col([](tCNode *c) { return c->getSatAnRatio(); }));
        w.f64("UnsatAnRatio", col([](tCNode *c) { return c->getUnsatAnRatio(); }));
        w.f64("Porosity", col([](tCNode *c) { return c->getPorosity(); }));
        w.f64("VolHeatCond", col([](tCNode *c) { return c->getVolHeatCond(); }));
        w.f64("SoilHeatCap", col([](tCNode *c) { return c->getSoilHeatCap(); }));
        w.f64("soil_cutoff", col([](tCNode *c) { return c->getSoilCutoff(); }));
        w.f64("root_cutoff", col([](tCNode *c) { return c->getRootCutoff(); }));
        // land
        w.f64("CanStorParam", col([](tCNode *c) { return c->getCanStorParam(); }));
        w.f64("IntercepCoeff", col([](tCNode *c) { return c->getIntercepCoeff(); }));
        w.f64("ThroughFall", col([](tCNode *c) { return c->getThroughFall(); }));
        w.f64("CanFieldCap", col([](tCNode *c) { return c->getCanFieldCap(); }));
        w.f64("DrainCoeff", col([](tCNode *c) { return c->getDrainCoeff(); }));
        w.f64("DrainExpPar", col([](tCNode *c) { return c->getDrainExpPar(); }));
        w.f64("LandUseAlb", col([](tCNode *c) { return c->getLandUseAlb(); }));
        w.f64("VegHeight", col([](tCNode *c) { return c->getVegHeight(); }));
        w.f64("OptTransmCoeff", col([](tCNode *c) { return c->getOptTransmCoeff(); }));
        w.f64("StomRes", col([](tCNode *c) { return c->getStomRes(); }));
        w.f64("VegFraction", col([](tCNode *c) { return c->getVegFraction(); }));
        w.f64("LeafAI", col([](tCNode *c) { return c->getLeafAI(); }));

        // Voronoi spokes in CCW order starting at getEdg() -> CSR over active cells.
        vector<int32_t> rowptr(N + 1, 0), nbr, eid, cid;
        vector<double> elen, evl, clen, cvl;
        // edge-list position of every active directed edge (list order = flux loop order,
        // reference: tHydroModel.cpp:2675)
        unordered_map<tEdge*, int> epos;
        {
            tMeshListIter<tEdge> eit(mesh->getEdgeList());
            tEdge *ce; int k = 0;
            for (ce = eit.FirstP(); eit.IsActive(); ce = eit.NextP()) epos[ce] = k++;
            w.i32s("n_active_edges", k);
        }
        vector<int32_t> elistpos;
        for (int i = 0; i < N; i++) {
            tEdge *first = cells[i]->getEdg();
            tEdge *e = first;
            int guard = 0;
            do {
                nbr.push_back(idx(e->getDestinationPtrNC()));
                eid.push_back(e->getID());
                elen.push_back(e->getLength());
                evl.push_back(e->getVEdgLen());
                auto ep = epos.find(e);
                elistpos.push_back(ep == epos.end() ? -1 : ep->second);
                tEdge *c = e->FindComplement();
                cid.push_back(c ? c->getID() : -1);
                clen.push_back(c ? c->getLength() : 0.0);
                cvl.push_back(c ? c->getVEdgLen() : 0.0);
                e = e->getCCWEdg();
            } while (e != first && ++guard < 1000);
            rowptr[i + 1] = (int32_t)nbr.size();
        }
        w.i32("spoke_rowptr", rowptr);
        w.i32("spoke_nbr", nbr);
        w.i32("spoke_edge_id", eid);
        w.i32("spoke_edge_listpos", elistpos);
        w.f64("spoke_len", elen);
        w.f64("spoke_vedglen", evl);
        w.i32("spoke_compl_id", cid);
        w.f64("spoke_compl_len", clen);
        w.f64("spoke_compl_vedglen", cvl);

        // active directed edge list (orig, dest) in list order
        {
            tMeshListIter<tEdge> eit(mesh->getEdgeList());
            tEdge *ce;
            vector<int32_t> eo, ed, eids;
            for (ce = eit.FirstP(); eit.IsActive(); ce = eit.NextP()) {
                eo.push_back(idx(ce->getOriginPtrNC()));
                ed.push_back(idx(ce->getDestinationPtrNC()));
                eids.push_back(ce->getID());
            }
            w.i32("edge_org", eo);
            w.i32("edge_dest", ed);
            w.i32("edge_id", eids);
        }

        // flow-net / routing scalars
        w.f64s("hillvel", flow->hillvel);
        w.f64s("streamvel", flow->streamvel);
        w.f64s("velratio", flow->velratio);
        w.f64s("velcoef", flow->velcoef);
        w.f64s("flowexp", flow->flowexp);
        w.f64s("baseflow", flow->baseflow);
        w.f64s("kincoef", flow->kincoef);
        w.f64s("dtReff", flow->dtReff);
        w.f64s("BasArea_flow", flow->BasArea);
        w.f64s("maxttime", flow->maxttime);
        w.i32s("outlet_stream_node", idx((tNode*)flow->OutletNode));
        w.f64s("outlet_contr_area", flow->OutletNode->getContrArea());
        w.i32s("res_limit", flow->res->limit);
        // reach lists
        {
            vector<int32_t> heads, outs, nn;
            tPtrListIter<tCNode> hi(flow->NodesLstH), oi(flow->NodesLstO);
            tCNode *c;
            for (c = hi.FirstP(); !hi.AtEnd(); c = hi.NextP()) heads.push_back(idx((tNode*)c));
            for (c = oi.FirstP(); !oi.AtEnd(); c = oi.NextP()) outs.push_back(idx((tNode*)c));
            tListNode<int> *p = flow->NNodes.getFirst();
            for (int k = 0; k < flow->NNodes.getSize(); k++, p = p->getNextNC()) nn.push_back(p->getDataNC());
            w.i32("reach_head", heads);
            w.i32("reach_outlet", outs);
            w.i32("reach_nnodes", nn);
        }
        // channel geometry and options of the kinematic-wave router (tKinemat.cpp:1136-1252 reads them per step)
        w.f64("chan_width", col([](tCNode *c) { return c->getChannelWidth(); }));
        w.f64s("outlet_chan_width", flow->OutletNode->getChannelWidth());
        w.f64s("chan_roughness", flow->Roughness);
        w.f64s("chan_width_uniform", flow->Width);
        w.i32s("chan_percolation", flow->percolationOption);
        w.i32s("chan_optres", flow->optres);
        // the reference's own reach partition for 2..8 partitions: tGraph::connectivity (tGraph.cpp:297-343) counts
        // the reaches of tKinemat's lists, tGraph::createDefaultPartition (588-608) assigns them; every cell then
        // belongs to reach2partition[tCNode::getReach()]. Pins tribs_b200/partition.py (tests/test_partition.py).
        if (graph_sim) {
            tGraph::sim = graph_sim; tGraph::mesh = mesh; tGraph::flow = flow;
            tGraph::connectivity();
            const int nr = tGraph::numGlobalReach;
            w.i32s("graph_n_reach", nr);
            for (int np = 2; np <= 8; np++) {
                if (nr < np) break;
                tGraph::reach2partition.assign(nr, 0);
                tGraph::createDefaultPartition(np);
                vector<int32_t> r2p(tGraph::reach2partition.begin(), tGraph::reach2partition.end());
                w.i32("graph_r2p_np" + std::to_string(np), r2p);
            }
            tGraph::conn.clear(); tGraph::reach2partition.clear(); tGraph::pointsPerReach.clear();
            delete[] tGraph::hid; delete[] tGraph::oid; tGraph::hid = tGraph::oid = nullptr;
            tGraph::numGlobalReach = 0;
        }
        // hydro model scalars / options
        w.f64s("BasArea", hydro->BasArea);
        w.f64s("IntStormMAX", hydro->IntStormMAX);
        w.f64s("surfaceSoilDepth", hydro->surfaceSoilDepth);
        w.f64s("rootZoneDepth", hydro->rootZoneDepth);
        w.i32s("EToption", hydro->EToption);
        w.i32s("Ioption", hydro->Ioption);
        w.i32s("SnOpt", hydro->SnOpt);
        w.i32s("RunOnoption", hydro->RunOnoption);
        w.i32s("RdstrOption", hydro->RdstrOption);
        w.i32s("GW_model_label", (int)hydro->simCtrl->GW_model_label);
        // timer
        w.f64s("tstep", timer->getTimeStep());
        w.f64s("gwstep", timer->getGWTimeStep());
        w.f64s("metstep", timer->getMetStep());
        w.f64s("etistep", timer->getEtIStep());
        w.f64s("endtime", timer->getEndTime());
        w.f64s("output_interval", timer->getOutputInterval());
        dump_et_setup(w);
        dump_state(w, "init_", true);
        w.f64("init_CanopyStorVol", col([](tCNode *c) { return c->getCanopyStorVol(); }));
        dump_et(w, "init_");
    }

    // -------------------------------------------------------------- ET / interception set-up
    // What tEvapoTrans / tIntercept read per cell or hold as members (tEvapoTrans.cpp:215-253,
    // 915-952, 3122-3330; tIntercept.cpp:44-45, 468-484): land classes, options, the weather
    // stations' hourly records and the Thiessen assignment of cells to stations.
    // with OPTSNOW the simulator drives the tSnowPack object (a tEvapoTrans) instead: its members count
    tEvapoTrans *active_et() { return (et->snowOption && snow) ? static_cast<tEvapoTrans *>(snow) : et; }

    void dump_et_setup(TrbWriter &w) {
        tEvapoTrans *E = active_et();
        w.i32s("et_option", E->evapotransOption);
        w.i32s("et_Ioption", E->Ioption);
        w.i32s("et_metdataOption", E->metdataOption);
        w.i32s("et_gFluxOption", E->gFluxOption);
        w.i32s("et_luOption", E->luOption);
        w.i32s("et_snowOption", E->snowOption);
        w.i32s("et_shelterOption", E->shelterOption);
        // tShelter's per-node horizon angle at azimuth 0: the only one tEvapoTrans::aboveHorizon ends up using
        w.f64("hor_angle_0000", col([](tCNode *c) { return c->getHorAngle0000(); }));
        w.f64s("et_timeStep", E->timeStep);
        w.f64s("et_Tlinke", E->Tlinke);
        w.f64s("et_tempLapseRate", E->tempLapseRate);
        w.f64s("icp_maxInterStormPeriod", icp->maxInterStormPeriod);
        w.f64s("sn_minSnTemp", snow->minSnTemp);
        w.f64s("sn_snliqfrac", snow->snliqfrac);
        w.i32s("sn_hillAlbedoOption", snow->hillAlbedoOption);
        w.i32s("icp_option", icp->interceptOption);
        vector<int32_t> lu(cells.size());
        int maxc = 0;
        for (size_t i = 0; i < cells.size(); i++) { lu[i] = cells[i]->getLandUse(); maxc = std::max(maxc, (int)lu[i]); }
        w.i32("land_class", lu);
        vector<double> table((size_t)(maxc + 1) * 15, 0.0);
        vector<char> seen(maxc + 1, 0);
        for (size_t i = 0; i < cells.size(); i++) {
            int c = lu[i];
            if (c < 0 || seen[c]) continue;
            seen[c] = 1;
            hydro->landPtr->setLandPtr(c);
            for (int k = 1; k <= 14; k++) table[(size_t)c * 15 + k] = hydro->landPtr->getLandProp(k);
        }
        w.f64("land_table", table);     // row = class id, column k = getLandProp(k), k = 1..14
        w.i32s("rain_type", rainfall->rainfallType);
        w.i32s("rain_optStorm", rainfall->getoptStorm());
        if (!rainfall->getoptStorm() && rainfall->rainfallType == 3) {   // gauges, tRainfall.cpp:463-503
            rainfall->assignStationToNode();   // what NewRainData does before every use (tRainfall.cpp:470)
            int ng = rainfall->numStations, T = rainfall->rainGauges[0].getTime();
            vector<int32_t> gid(ng), asg(cells.size());
            vector<double> gel(ng), ser((size_t)ng * T);
            for (int k = 0; k < ng; k++) {
                gid[k] = rainfall->rainGauges[k].getStation(); gel[k] = rainfall->rainGauges[k].getElev();
                for (int t = 0; t < T; t++) ser[(size_t)k * T + t] = rainfall->rainGauges[k].getRain(t);
            }
            for (size_t i = 0; i < cells.size(); i++) asg[i] = rainfall->assignedRain[i];
            w.i32("rain_gauge_id", gid); w.f64("rain_gauge_elev", gel); w.f64("rain_series", ser);
            w.i32("rain_assigned_gauge", asg); w.i32s("rain_n_times", T);
            w.f64s("rain_precLapseRate", rainfall->precLapseRate);
            w.f64s("rain_dt", timer->getRainDT());
        }
        if (E->evapotransOption == 0 || E->metdataOption != 1 || E->weatherStations == nullptr) return;
        int ns = E->numStations;
        w.i32s("met_n_stations", ns);
        vector<int32_t> sid(ns), gmt(ns), nt(ns), np(ns);
        vector<double> lat(ns), lon(ns), oth(ns);
        for (int k = 0; k < ns; k++) {
            tHydroMet &st = E->weatherStations[k];
            sid[k] = st.getStation(); gmt[k] = st.getGmt(); nt[k] = st.getTime(); np[k] = st.getParm();
            lat[k] = st.getLat(1); lon[k] = st.getLong(1); oth[k] = st.getOther();
        }
        w.i32("met_station_id", sid); w.i32("met_gmt", gmt); w.i32("met_n_times", nt); w.i32("met_n_params", np);
        w.f64("met_lat", lat); w.f64("met_long", lon); w.f64("met_other", oth);
        int T = nt[0];
        auto series = [&](const char *name, double (tHydroMet::*get)(int)) {
            vector<double> v((size_t)ns * T);
            for (int k = 0; k < ns; k++) for (int t = 0; t < T; t++) v[(size_t)k * T + t] = (E->weatherStations[k].*get)(t);
            w.f64(name, v);
        };
        series("met_airTemp", &tHydroMet::getAirTemp);
        series("met_dewTemp", &tHydroMet::getDewTemp);
        series("met_surfTemp", &tHydroMet::getSurfTemp);
        series("met_rHumidity", &tHydroMet::getRHumidity);
        series("met_vPress", &tHydroMet::getVaporPress);
        series("met_atmPress", &tHydroMet::getAtmPress);
        series("met_windSpeed", &tHydroMet::getWindSpeed);
        series("met_skyCover", &tHydroMet::getSkyCover);
        series("met_radGlobal", &tHydroMet::getRadGlobal);
        vector<int32_t> cal((size_t)T * 4);
        for (int t = 0; t < T; t++) {
            tHydroMet &st = E->weatherStations[0];
            cal[t * 4 + 0] = st.getYear(t); cal[t * 4 + 1] = st.getMonth(t);
            cal[t * 4 + 2] = st.getDay(t); cal[t * 4 + 3] = st.getHour(t);
        }
        w.i32("met_calendar", cal);
        vector<int32_t> as(cells.size());
        for (size_t i = 0; i < cells.size(); i++) as[i] = E->assignedStation[i];
        w.i32("met_assigned_station", as);
        w.i32s("met_tsOption", E->tsOption);
        w.i32s("met_vapOption", E->vapOption);
        vector<int32_t> start = {timer->year, timer->month, timer->day, timer->hour, timer->minute};
        w.i32("timer_start", start);
    }

    // per-cell outputs of tEvapoTrans::callEvapoPotential / callEvapoTrans and tIntercept, plus
    // the members the host-side mirror has to reproduce (sun geometry, met clock)
    void dump_et(TrbWriter &w, const string &pre) {
        if (!et || et->evapotransOption == 0) return;
        tEvapoTrans *E = active_et();
        w.f64(pre + "PotEvap", col([](tCNode *c) { return c->getPotEvap(); }));
        w.f64(pre + "ActEvap", col([](tCNode *c) { return c->getActEvap(); }));
        w.f64(pre + "SurfTemp", col([](tCNode *c) { return c->getSurfTemp(); }));
        w.f64(pre + "SoilTemp", col([](tCNode *c) { return c->getSoilTemp(); }));
        w.f64(pre + "NetRad", col([](tCNode *c) { return c->getNetRad(); }));
        w.f64(pre + "GFlux", col([](tCNode *c) { return c->getGFlux(); }));
        w.f64(pre + "HFlux", col([](tCNode *c) { return c->getHFlux(); }));
        w.f64(pre + "LFlux", col([](tCNode *c) { return c->getLFlux(); }));
        w.f64(pre + "LongRadIn", col([](tCNode *c) { return c->getLongRadIn(); }));
        w.f64(pre + "LongRadOut", col([](tCNode *c) { return c->getLongRadOut(); }));
        w.f64(pre + "ShortRadIn", col([](tCNode *c) { return c->getShortRadIn(); }));
        w.f64(pre + "ShortRadIn_dir", col([](tCNode *c) { return c->getShortRadIn_dir(); }));
        w.f64(pre + "ShortRadIn_dif", col([](tCNode *c) { return c->getShortRadIn_dif(); }));
        w.f64(pre + "ShortRadSlope", col([](tCNode *c) { return c->getShortRadSlope(); }));
        w.f64(pre + "ShortAbsbVeg", col([](tCNode *c) { return c->getShortAbsbVeg(); }));
        w.f64(pre + "ShortAbsbSoi", col([](tCNode *c) { return c->getShortAbsbSoi(); }));
        w.f64(pre + "AirTemp", col([](tCNode *c) { return c->getAirTemp(); }));
        w.f64(pre + "DewTemp", col([](tCNode *c) { return c->getDewTemp(); }));
        w.f64(pre + "RelHumid", col([](tCNode *c) { return c->getRelHumid(); }));
        w.f64(pre + "VapPressure", col([](tCNode *c) { return c->getVapPressure(); }));
        w.f64(pre + "SkyCover", col([](tCNode *c) { return c->getSkyCover(); }));
        w.f64(pre + "WindSpeed", col([](tCNode *c) { return c->getWindSpeed(); }));
        w.f64(pre + "AirPressure", col([](tCNode *c) { return c->getAirPressure(); }));
        w.f64(pre + "CumHrsSun", col([](tCNode *c) { return c->getCumHrsSun(); }));
        w.f64(pre + "AvEvapFract", col([](tCNode *c) { return c->getAvEvapFract(); }));
        w.f64(pre + "SheltFact", col([](tCNode *c) { return c->getSheltFact(); }));
        w.f64(pre + "CumRSin", col([](tCNode *c) { return c->getCumRSin(); }));
        w.f64(pre + "EvapWetCanopy", col([](tCNode *c) { return c->getEvapWetCanopy(); }));
        w.f64(pre + "CumTotEvap", col([](tCNode *c) { return c->getCumTotEvap(); }));
        w.f64(pre + "CumBarEvap", col([](tCNode *c) { return c->getCumBarEvap(); }));
        w.f64(pre + "AvET", col([](tCNode *c) { return c->getAvET(); }));
        w.f64(pre + "InterceptLoss", col([](tCNode *c) { return c->getInterceptLoss(); }));
        w.f64(pre + "CanStorage", col([](tCNode *c) { return c->getCanStorage(); }));
        w.f64(pre + "CumIntercept", col([](tCNode *c) { return c->getCumIntercept(); }));
        w.f64(pre + "StormLength", col([](tCNode *c) { return c->getStormLength(); }));
        w.f64(pre + "EvapoTrans", col([](tCNode *c) { return c->getEvapoTrans(); }));
        w.f64(pre + "EvapSoil", col([](tCNode *c) { return c->getEvapSoil(); }));
        w.f64(pre + "EvapDryCanopy", col([](tCNode *c) { return c->getEvapDryCanopy(); }));
        w.f64(pre + "NetPrecipitation", col([](tCNode *c) { return c->getNetPrecipitation(); }));
        w.f64(pre + "Rain", col([](tCNode *c) { return c->getRain(); }));
        if (et->snowOption) {   // tSnowPack state, fluxes and integrated outputs (tSnowPack.cpp:1052-1135)
#define SNF(name, getter) w.f64(pre + name, col([](tCNode *c) { return c->getter(); }))
            SNF("LiqWE", getLiqWE); SNF("IceWE", getIceWE); SNF("SnTempC", getSnTempC); SNF("LiqRouted", getLiqRouted);
            SNF("CrustAge", getCrustAge); SNF("DensityAge", getDensityAge); SNF("EvapoTransAge", getEvapoTransAge);
            SNF("SnSub", getSnSub); SNF("SnEvap", getSnEvap); SNF("DU", getDU); SNF("Unode", getUnode); SNF("Uerror", getUerror);
            SNF("SnLHF", getSnLHF); SNF("SnSHF", getSnSHF); SNF("SnGHF", getSnGHF); SNF("SnPHF", getSnPHF);
            SNF("SnRLout", getSnRLout); SNF("SnRLin", getSnRLin); SNF("SnRSin", getSnRSin);
            SNF("IntSWE", getIntSWE); SNF("IntPrec", getIntPrec); SNF("IntSnUnload", getIntSnUnload); SNF("IntSub", getIntSub);
            SNF("CumIntSub", getCumIntSub); SNF("CumIntUnl", getCumIntUnl);
            SNF("CumLHF", getCumLHF); SNF("CumMelt", getCumMelt); SNF("CumSHF", getCumSHF); SNF("CumSnSub", getCumSnSub);
            SNF("CumSnEvap", getCumSnEvap); SNF("CumPHF", getCumPHF); SNF("CumRLin", getCumRLin); SNF("CumRLout", getCumRLout);
            SNF("CumGHF", getCumGHF); SNF("CumUerror", getCumUerror); SNF("CumHrsSnow", getCumHrsSnow);
            SNF("PersTimeMax", getPersTimeMax); SNF("PersTime", getPersTime); SNF("PeakSWE", getPeakSWE);
            SNF("PeakSWETemp", getPeakSWETemp); SNF("InitPackTime", getInitPackTime); SNF("InitPackTimeTemp", getInitPackTimeTemp);
            SNF("PeakPackTime", getPeakPackTime);
#undef SNF
        }
        vector<double> sun = {E->alphaD, E->sinAlpha, E->sunaz, E->Io, (double)E->currentTime[3],
                              (double)E->hourlyTimeStep, (double)E->dewHumFlag, (double)E->skycover_flag,
                              E->del, E->phi, E->tau, E->deltaT, (double)E->nodeHour,
                              E->latitude, E->longitude, (double)E->gmt};
        w.f64(pre + "sun", sun);
    }

    // -------------------------------------------------------------- dynamic state
    void dump_state(TrbWriter &w, const string &pre, bool with_old) {
        w.f64(pre + "NwtNew", col([](tCNode *c) { return c->getNwtNew(); }));
        w.f64(pre + "MuNew", col([](tCNode *c) { return c->getMuNew(); }));
        w.f64(pre + "MiNew", col([](tCNode *c) { return c->getMiNew(); }));
        w.f64(pre + "NtNew", col([](tCNode *c) { return c->getNtNew(); }));
        w.f64(pre + "NfNew", col([](tCNode *c) { return c->getNfNew(); }));
        w.f64(pre + "RuNew", col([](tCNode *c) { return c->getRuNew(); }));
        w.f64(pre + "RiNew", col([](tCNode *c) { return c->getRiNew(); }));
        if (with_old) {
            w.f64(pre + "Hlevel", col([](tCNode *c) { return c->getHlevel(); }));      // channel router (tKinemat)
            w.f64(pre + "Qstrm", col([](tCNode *c) { return c->getQstrm(); }));
            w.f64(pre + "NwtOld", col([](tCNode *c) { return c->getNwtOld(); }));
            w.f64(pre + "MuOld", col([](tCNode *c) { return c->getMuOld(); }));
            w.f64(pre + "MiOld", col([](tCNode *c) { return c->getMiOld(); }));
            w.f64(pre + "NtOld", col([](tCNode *c) { return c->getNtOld(); }));
            w.f64(pre + "NfOld", col([](tCNode *c) { return c->getNfOld(); }));
            w.f64(pre + "RuOld", col([](tCNode *c) { return c->getRuOld(); }));
            w.f64(pre + "RiOld", col([](tCNode *c) { return c->getRiOld(); }));
        }
        w.f64(pre + "Qpin", col([](tCNode *c) { return c->getQpin(); }));
        w.f64(pre + "Qpout", col([](tCNode *c) { return c->getQpout(); }));
        w.f64(pre + "intstorm", col([](tCNode *c) { return c->getIntStormVar(); }));
        w.f64(pre + "gwchange", col([](tCNode *c) { return c->gwchange; }));
        w.f64(pre + "srf", col([](tCNode *c) { return c->getSrf(); }));
        w.f64(pre + "hsrf", col([](tCNode *c) { return c->getHsrf(); }));
        w.f64(pre + "sbsrf", col([](tCNode *c) { return c->getSbsrf(); }));
        w.f64(pre + "psrf", col([](tCNode *c) { return c->getPsrf(); }));
        w.f64(pre + "satsrf", col([](tCNode *c) { return c->getSatsrf(); }));
        w.f64(pre + "esrf", col([](tCNode *c) { return c->getEsrf(); }));
        w.f64(pre + "srf_hr", col([](tCNode *c) { return c->getSrf_Hr(); }));
        w.f64(pre + "cumsrf", col([](tCNode *c) { return c->getCumSrf(); }));
        w.f64(pre + "Rain", col([](tCNode *c) { return c->getRain(); }));
        w.f64(pre + "SoilMoisture", col([](tCNode *c) { return c->getSoilMoisture(); }));
        w.f64(pre + "SoilMoistureSC", col([](tCNode *c) { return c->getSoilMoistureSC(); }));
        w.f64(pre + "SoilMoistureUNSC", col([](tCNode *c) { return c->getSoilMoistureUNSC(); }));
        w.f64(pre + "RootMoisture", col([](tCNode *c) { return c->getRootMoisture(); }));
        w.f64(pre + "RootMoistureSC", col([](tCNode *c) { return c->getRootMoistureSC(); }));
        w.f64(pre + "AvSoilMoisture", col([](tCNode *c) { return c->getAvSoilMoisture(); }));
        w.f64(pre + "Recharge", col([](tCNode *c) { return c->getRecharge(); }));
        w.f64(pre + "UnSatFlowIn", col([](tCNode *c) { return c->getUnSatFlowIn(); }));
        w.f64(pre + "UnSatFlowOut", col([](tCNode *c) { return c->getUnSatFlowOut(); }));
        w.f64(pre + "Transmissivity", col([](tCNode *c) { return c->getTransmiss(); }));
        w.f64(pre + "hsrfOccur", col([](tCNode *c) { return c->hsrfOccur; }));
        w.f64(pre + "sbsrfOccur", col([](tCNode *c) { return c->sbsrfOccur; }));
        w.f64(pre + "psrfOccur", col([](tCNode *c) { return c->psrfOccur; }));
        w.f64(pre + "satsrfOccur", col([](tCNode *c) { return c->satsrfOccur; }));
        w.f64(pre + "satOccur", col([](tCNode *c) { return (double)c->satOccur; }));
        w.f64(pre + "RechDisch", col([](tCNode *c) { return c->RechDisch; }));
        w.f64(pre + "FlowVelocity", col([](tCNode *c) { return c->getFlowVelocity(); }));
        w.f64(pre + "TTime", col([](tCNode *c) { return c->getTTime(); }));
        w.f64(pre + "Qstrm", col([](tCNode *c) { return c->getQstrm(); }));
        w.f64(pre + "Hlevel", col([](tCNode *c) { return c->getHlevel(); }));
        w.f64(pre + "NetPrecipitation", col([](tCNode *c) { return c->getNetPrecipitation(); }));
        w.f64(pre + "EvapSoil", col([](tCNode *c) { return c->getEvapSoil(); }));
        w.f64(pre + "EvapDryCanopy", col([](tCNode *c) { return c->getEvapDryCanopy(); }));
        w.f64(pre + "EvapoTrans", col([](tCNode *c) { return c->getEvapoTrans(); }));
        w.f64(pre + "UnSaturatedStorage", col([](tCNode *c) { return c->getUnSaturatedStorage(); }));
        w.f64(pre + "SaturatedStorage", col([](tCNode *c) { return c->getSaturatedStorage(); }));
        // basin accumulators of tHydroModel
        vector<double> g = {hydro->Stok, hydro->TotRain, hydro->TotGWchange, hydro->TotMoist,
                            hydro->psrf, timer->getCurrentTime(), flow->Qout,
                            hydro->fSoi100, hydro->fTop100, hydro->fClm100, hydro->fGW100,
                            hydro->dM100, hydro->dMRt, hydro->mTh100, hydro->mThRt};
        w.f64(pre + "globals", g);
    }

    void dump_routing(TrbWriter &w, const string &pre) {
        {   // channel router state that is not a per-cell field: tKinemat::OutletHlev and the basin outlet node
            vector<double> oh(flow->OutletHlev, flow->OutletHlev + flow->NodesLstO.getSize());
            w.f64(pre + "OutletHlev", oh);
            w.f64s(pre + "outlet_Hlevel", flow->OutletNode->getHlevel());
            w.f64s(pre + "outlet_Qstrm", flow->OutletNode->getQstrm());
        }
        // (TimeInd, Qeff) stacks of the stream cells (reference: tKinemat.cpp:1023-1067)
        vector<int32_t> off(1, 0), node, tind;
        vector<double> q;
        for (size_t i = 0; i < cells.size(); i++) {
            tCNode *c = cells[i];
            if (c->getBoundaryFlag() != kStream) continue;
            tList<int> *ti = c->getTimeIndList();
            tList<double> *qe = c->getQeffList();
            if (!ti || !qe) continue;
            node.push_back((int)i);
            tListNode<int> *a = ti->getFirst();
            tListNode<double> *b = qe->getFirst();
            for (int k = 0; k < ti->getSize(); k++, a = a->getNextNC(), b = b->getNextNC()) {
                tind.push_back(a->getDataNC());
                q.push_back(b->getDataNC());
            }
            off.push_back((int32_t)tind.size());
        }
        w.i32(pre + "qeff_node", node);
        w.i32(pre + "qeff_off", off);
        w.i32(pre + "qeff_tind", tind);
        w.f64(pre + "qeff_val", q);
        tFlowResults *r = flow->res;
        int lim = r->limit;
        auto arr = [&](const char *nm, double *p) {
            w.rec(pre + nm, 0, lim, p, 8);
        };
        arr("mhydro", r->mhydro);
        arr("HsrfRout", r->HsrfRout);
        arr("SbsrfRout", r->SbsrfRout);
        arr("PsrfRout", r->PsrfRout);
        arr("SatsrfRout", r->SatsrfRout);
        arr("crr", r->crr);
        arr("rmax", r->max);
        arr("rmin", r->min);
        arr("frac", r->frac);
        arr("msm", r->msm);
        arr("msmRt", r->msmRt);
        arr("msmU", r->msmU);
        arr("mgw", r->mgw);
        arr("sat", r->sat);
        arr("qunsat", r->qunsat);
    }

    // -------------------------------------------------------------- known answers
    void dump_kat(const string &path) {
        TrbWriter w(path);
        tCNode *cn = cells[0];
        hydro->SetupNodeUSZ(cn);
        hydro->DtoBedrock = cn->getBedrockDepth();
        vector<double> prm = {hydro->Ksat, hydro->Ths, hydro->Thr, hydro->PoreInd, hydro->Psib,
                              hydro->F, hydro->Ar, hydro->UAr, hydro->Eps, hydro->DtoBedrock};
        w.f64("kat_soil", prm);
        // LambertW (reference: tHydroModel.cpp:3831-3915)
        {
            vector<double> zi, zo;
            for (int k = 0; k <= 400; k++) {
                double z = -0.3678 + 0.3678 * k / 400.0 * 0.999999;
                zi.push_back(z);
            }
            for (int k = 1; k <= 60; k++) zi.push_back(-pow(10.0, -0.25 * k));
            for (int k = 1; k <= 40; k++) zi.push_back(0.05 * k);
            zi.push_back(1.0); zi.push_back(2.5); zi.push_back(10.0); zi.push_back(0.0);
            for (double z : zi) zo.push_back(hydro->LambertW(z));
            w.f64("lambertw_in", zi);
            w.f64("lambertw_out", zo);
        }
        // get_Total_Moist / get_Upper_Moist / get_Lower_Moist
        {
            vector<double> a, o;
            for (int k = 0; k <= 300; k++) {
                double nwt = 50.0 * k;
                a.push_back(nwt);
                double save = hydro->NwtNew;
                o.push_back(hydro->get_Total_Moist(nwt));
                hydro->NwtNew = save;
            }
            w.f64("total_moist_in", a);
            w.f64("total_moist_out", o);
            vector<double> nf, nw, up, lo;
            for (int i = 1; i <= 40; i++)
                for (int j = 1; j <= 12; j++) {
                    double nwt = 250.0 + 1200.0 * j;
                    double f = nwt * i / 41.0;
                    nf.push_back(f); nw.push_back(nwt);
                    up.push_back(hydro->get_Upper_Moist(f, nwt));
                    lo.push_back(hydro->get_Lower_Moist(f, nwt));
                }
            w.f64("ul_moist_nf", nf);
            w.f64("ul_moist_nwt", nw);
            w.f64("upper_moist_out", up);
            w.f64("lower_moist_out", lo);
        }
        // Newton #1 (water-table solve, reference: tHydroModel.cpp:3937-4020)
        {
            vector<double> dm, nw, o;
            for (int i = 0; i <= 30; i++)
                for (int j = 0; j <= 10; j++) {
                    double nwt = (j == 0) ? 0.0 : 300.0 + 1400.0 * j;
                    double d = 5.0 + 40.0 * i * (1 + j);
                    dm.push_back(d); nw.push_back(nwt);
                    double save = hydro->NwtNew;
                    o.push_back(hydro->Newton(d, nwt));
                    hydro->NwtNew = save;
                }
            w.f64("newton1_dm", dm);
            w.f64("newton1_nwt", nw);
            w.f64("newton1_out", o);
        }
        // Newton #2 (front relocation, reference: tHydroModel.cpp:4037-4082)
        {
            vector<double> dm, nw, nf, fl, o;
            for (int flag = 0; flag <= 1; flag++)
                for (int i = 1; i <= 12; i++)
                    for (int j = 1; j <= 8; j++) {
                        double nwt = 1500.0 + 1000.0 * j;
                        double f = (nwt - 250.0) * i / 14.0;
                        double d = 0.2 * i + 0.05 * j;
                        dm.push_back(d); nw.push_back(nwt); nf.push_back(f); fl.push_back(flag);
                        o.push_back(hydro->Newton(d, nwt, f, flag));
                    }
            w.f64("newton2_dm", dm);
            w.f64("newton2_nwt", nw);
            w.f64("newton2_nf", nf);
            w.f64("newton2_flag", fl);
            w.f64("newton2_out", o);
        }
        // get_RechargeRate, lateral flows, Z1Z2
        {
            vector<double> a, b, o, us, ss, zz;
            for (int i = 1; i <= 30; i++)
                for (int j = 1; j <= 10; j++) {
                    double nf = 20.0 * i;
                    double dmv = hydro->Thr * nf + (hydro->Ths - hydro->Thr) * nf * j / 11.0;
                    a.push_back(dmv); b.push_back(nf);
                    double ru = hydro->get_RechargeRate(dmv, nf);
                    o.push_back(ru);
                    us.push_back(hydro->get_UnSat_LateralFlow(nf, ru, 30000.0));
                    ss.push_back(hydro->get_Sat_LateralFlow(nf * j / 11.0 * (j % 3 ? 1 : 0), nf, ru, 30000.0));
                    zz.push_back(hydro->get_Z1Z2_Moist(nf, nf + 100.0 * j, 3000.0, 0));
                }
            w.f64("recharge_dm", a);
            w.f64("recharge_nf", b);
            w.f64("recharge_out", o);
            w.f64("unsat_lat_out", us);
            w.f64("sat_lat_out", ss);
            w.f64("z1z2_out", zz);
        }
    }
};

// ---------------------------------------------------------------------------------
// --dump-meshb DIR: the mesh and flow network the reference has just built, in the OPTMESHINPUT 9
// ("MeshBuilder") binary formats its own readers define (tMesh::MakeMeshFromMeshBuilder,
// tMesh.cpp:1076-1306; tFlowNet::ReadFlowNetFromMeshBuilder, tFlowNet.cpp:117-175). Used to check
// tribs_b200/meshb.py (which writes the same files for the large synthetic basins) against the
// reference: a deck re-run with OPTMESHINPUT 9 on these files must reproduce the run bit for bit.
template <class T> static void bput(ofstream &f, T v) { f.write((const char *)&v, sizeof(T)); }

static void dump_meshb(const string &dir, tMesh<tCNode> &mesh, tKinemat &flow) {
    ofstream nf(dir + "/nodes.meshb", ios::binary), ef(dir + "/edges.meshb", ios::binary), ff(dir + "/flow.meshb", ios::binary);
    tMeshListIter<tCNode> ni(mesh.getNodeList());
    bput<int>(nf, mesh.getNodeList()->getSize());
    for (tCNode *c = ni.FirstP(); !ni.AtEnd(); c = ni.NextP()) {
        bput<int>(nf, c->getID()); bput<int>(nf, c->getBoundaryFlag());
        bput<double>(nf, c->getX()); bput<double>(nf, c->getY()); bput<double>(nf, c->getZ()); bput<double>(nf, c->getVArea());
        bput<int>(nf, c->getEdg()->getID()); bput<int>(nf, c->getFloodStatus()); bput<int>(nf, c->getTracer());
        bput<double>(nf, c->getHillPath()); bput<double>(nf, c->getTTime());
        bput<double>(nf, c->getSrf()); bput<double>(nf, c->getHsrf()); bput<double>(nf, c->getPsrf());
        bput<double>(nf, c->getSatsrf()); bput<double>(nf, c->getSbsrf()); bput<double>(nf, c->getEvapoTrans());
        bput<double>(nf, c->getSoilMoistureSC()); bput<double>(nf, c->getSoilMoistureUNSC()); bput<double>(nf, c->getRootMoistureSC());
        bput<double>(nf, c->getContrArea()); bput<double>(nf, c->getNwtNew()); bput<double>(nf, c->getRain());
        bput<double>(nf, c->getCurvature()); bput<double>(nf, c->getStreamPath());
        bput<int>(nf, c->getReach());
        bput<int>(nf, c->getFlowEdg() ? c->getFlowEdg()->getID() : -1);
        bput<int>(nf, c->getStreamNode() ? c->getStreamNode()->getID() : -1);
    }
    tMeshListIter<tEdge> ei(mesh.getEdgeList());
    bput<int>(ef, mesh.getEdgeList()->getSize());
    for (tEdge *e = ei.FirstP(); !ei.AtEnd(); e = ei.NextP()) {
        tArray<double> rv = e->getRVtx();
        bput<int>(ef, e->getID()); bput<double>(ef, rv[0]); bput<double>(ef, rv[1]);
        bput<double>(ef, e->getLength()); bput<double>(ef, e->getSlope()); bput<double>(ef, e->getVEdgLen());
        bput<int>(ef, e->getOriginPtrNC()->getID()); bput<int>(ef, e->getDestinationPtrNC()->getID());
        bput<int>(ef, e->getCCWEdg()->getID());
    }
    bput<double>(ff, flow.hillvel); bput<double>(ff, flow.streamvel); bput<double>(ff, flow.velratio); bput<double>(ff, flow.velcoef);
    bput<double>(ff, flow.flowexp); bput<double>(ff, flow.baseflow); bput<double>(ff, flow.flowout); bput<double>(ff, flow.maxttime);
    bput<double>(ff, flow.dist_hill_max); bput<double>(ff, flow.dist_stream_max); bput<double>(ff, flow.BasArea);
    bput<int>(ff, flow.OutletNode->getID());
    tPtrList<tCNode> *lists[3] = {&flow.HeadsLst, &flow.NodesLstH, &flow.NodesLstO};
    for (auto *L : lists) {
        tPtrListIter<tCNode> it(*L);
        bput<int>(ff, L->getSize());
        for (tCNode *c = it.FirstP(); !it.AtEnd(); c = it.NextP()) bput<int>(ff, c->getID());
    }
    tListIter<int> ci(flow.NNodes);
    bput<int>(ff, flow.NNodes.getSize());
    for (int *v = ci.FirstP(); !ci.AtEnd(); v = ci.NextP()) bput<int>(ff, *v);
}

static double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    // harness flags are stripped before the reference's own CLI parser sees argv
    string outdir = ".";
    int snap_from = 1, snap_to = 0, snap_every = 1;
    string phases = "ugre";
    bool do_kat = false, do_basin = true;
    int max_steps = 0;   // --max-steps N: leave the loop after N steps (RUNTIME still sizes the hydrograph arrays)
    int time_from = 0;   // --time-from W: the first W steps are warm-up, left out of the timing line
    string meshb_dir; // --dump-meshb <dir>: write the OPTMESHINPUT 9 files of the mesh / flow network just built
    string gpu_lib;   // --gpu <libtribs_gpu.so>: the reference's loop drives the B200 path
    int gpu_ranks = 1, gpu_batch = 16;
    bool gpu_channels = false;
    double inject_evap = 0.0;   // --inject-evap R: soil evaporation R mm/h on dry steps (see below)
    vector<char*> rargs;
    for (int i = 0; i < argc; i++) {
        string a = argv[i];
        if (a == "--out" && i + 1 < argc) outdir = argv[++i];
        else if (a == "--snap-from" && i + 1 < argc) snap_from = atoi(argv[++i]);
        else if (a == "--snap-to" && i + 1 < argc) snap_to = atoi(argv[++i]);
        else if (a == "--snap-every" && i + 1 < argc) snap_every = atoi(argv[++i]);
        else if (a == "--phases" && i + 1 < argc) phases = argv[++i];
        else if (a == "--kat") do_kat = true;
        else if (a == "--gpu" && i + 1 < argc) gpu_lib = argv[++i];
        else if (a == "--gpu-ranks" && i + 1 < argc) gpu_ranks = atoi(argv[++i]);     // reach partitions, one handle each
        else if (a == "--gpu-batch" && i + 1 < argc) gpu_batch = atoi(argv[++i]);     // steps per tribs_gpu_run
        else if (a == "--gpu-channels") gpu_channels = true;    // partitioned mode: kinematic-wave reaches on the device too
        else if (a == "--inject-evap" && i + 1 < argc) inject_evap = atof(argv[++i]);
        else if (a == "--no-basin") do_basin = false;
        else if (a == "--dump-meshb" && i + 1 < argc) meshb_dir = argv[++i];
        else if (a == "--max-steps" && i + 1 < argc) max_steps = atoi(argv[++i]);
        else if (a == "--time-from" && i + 1 < argc) time_from = atoi(argv[++i]);
        else rargs.push_back(argv[i]);
    }
    int rargc = (int)rargs.size();
    char **rargv = rargs.data();

    double t_setup0 = now_s();
    SimulationControl SimCtrl(rargc, rargv);
    tInputFile InputFile(SimCtrl.infile);
    tPreProcess PreProcessor(&SimCtrl, InputFile);
    tRunTimer Timer(InputFile);
    tMesh<tCNode> BasinMesh(&SimCtrl, InputFile);
    tKinemat Flow(&SimCtrl, &BasinMesh, InputFile, &Timer);
    if (!meshb_dir.empty()) dump_meshb(meshb_dir, BasinMesh, Flow);
    tResample RsmplMaster(&SimCtrl, &BasinMesh);
    tShelter Shelter(&SimCtrl, &BasinMesh, InputFile);
    tCOutput<tCNode> Output(&SimCtrl, &BasinMesh, InputFile, &RsmplMaster, &Timer);
    tWaterBalance Balance(&SimCtrl, &BasinMesh, InputFile);
    tHydroModel Moisture(&SimCtrl, &BasinMesh, InputFile, &RsmplMaster, &Balance, &Timer);
    tRainfall Rainfall(&SimCtrl, &BasinMesh, InputFile, &RsmplMaster);
    tEvapoTrans EvapoTrans(&SimCtrl, &BasinMesh, InputFile, &Timer, &RsmplMaster, &Moisture, &Rainfall);
    tIntercept Intercept(&SimCtrl, &BasinMesh, InputFile, &Timer, &RsmplMaster, &Moisture);
    tSnowPack SnowPack(&SimCtrl, &BasinMesh, InputFile, &Timer, &RsmplMaster, &Moisture, &Rainfall);
    tRestart<tCNode> Restart(&Timer, &BasinMesh, &Flow, &Balance, &Moisture, &Rainfall,
                             &EvapoTrans, &Intercept, &SnowPack);
    Simulator Sim(&SimCtrl, &Rainfall, &Timer, &Output, &Restart);
    Sim.initialize_simulation(&EvapoTrans, &SnowPack, InputFile);
    double t_setup1 = now_s();

    // Net evaporation without the ET module: the unsaturated zone only sees ET through the node
    // fields EvapSoil / EvapDryCanopy (tHydroModel.cpp:795-809). Switching its EToption on and
    // writing EvapSoil here drives the reference's own code into the states that need R1 < 0
    // (WTDropsFromSurf, IntStormBelow) with a forcing that is trivial to reproduce elsewhere.
    if (inject_evap > 0.0) Moisture.EToption = 1;

    Harness H;
    H.mesh = &BasinMesh; H.flow = &Flow; H.hydro = &Moisture; H.timer = &Timer;
    H.et = &EvapoTrans; H.icp = &Intercept; H.rainfall = &Rainfall; H.snow = &SnowPack;
    H.index();
    H.graph_sim = &SimCtrl;
    if (do_basin) H.dump_basin(outdir + "/basin.trb", InputFile);
    if (do_kat) H.dump_kat(outdir + "/kat.trb");

    TribsGpuBinding gpu;
    const bool use_gpu = !gpu_lib.empty();
    vector<int32_t> gpu_owner;
    if (use_gpu && gpu_ranks > 1) {
        // the reference's own partition: tGraph::connectivity + createDefaultPartition on tKinemat's reach lists,
        // every cell with reach2partition[tCNode::getReach()] (what tGraph::update / inLocalPartition use)
        tGraph::sim = &SimCtrl; tGraph::mesh = &BasinMesh; tGraph::flow = &Flow;
        tGraph::connectivity();
        const int nr = tGraph::numGlobalReach;
        if (nr < gpu_ranks) { cout << "\nError: " << nr << " reaches cannot be split over " << gpu_ranks << " partitions" << endl; exit(2); }
        tGraph::reach2partition.assign(nr, 0);
        tGraph::createDefaultPartition(gpu_ranks);
        for (tCNode *cn : H.cells) gpu_owner.push_back(tGraph::reach2partition[cn->getReach()]);
    }
    if (use_gpu) gpu.attach(&BasinMesh, &Flow, &Moisture, &Timer, gpu_lib.c_str(), gpu_ranks, &gpu_owner, gpu_batch);
    const bool gpu_et = use_gpu && EvapoTrans.getEToption() != 0 && SnowPack.getSnowOpt() == 0;
    if (gpu_et) gpu.attachET(&EvapoTrans, &Intercept, &Moisture);
    const bool gpu_snow = use_gpu && EvapoTrans.getEToption() != 0 && SnowPack.getSnowOpt() != 0;
    if (gpu_snow) gpu.attachSnow(&SnowPack, &Intercept, &Moisture);
    vector<double> qout_series;

    std::unique_ptr<TrbWriter> sw;
    if (snap_to >= snap_from) sw.reset(new TrbWriter(outdir + "/steps.trb"));

    // Partitioned hybrid run: the reference's timer and rainfall objects advance k steps ahead, the device runs them as
    // ONE pipelined batch on every rank (tribs_gpu_run, peer-memory halos), the owned cells of every rank go back to
    // the tCNodes. With --gpu-channels the reaches are routed on the device as well (tribs_gpu_channel_route over the
    // batch) and tKinemat's levels and discharges are refilled from it; snapshots are taken after Reset().
    if (use_gpu && gpu_ranks > 1) {
        int step = 0;
        if (gpu_channels) gpu.attachChannels(&Flow);
        vector<int32_t> snap_steps;
        while (!Timer.IsFinished() && !(max_steps > 0 && step >= max_steps)) {
            const int k_max = max_steps > 0 ? std::min(gpu_batch, max_steps - step) : gpu_batch;
            vector<tribs_step_t> st;
            vector<uint32_t> masks;
            vector<vector<double>> rains;
            for (int k = 0; k < k_max && !Timer.IsFinished(); k++) {
                Timer.Advance(Timer.getTimeStep());
                Sim.UpdatePrecipitationInput(Rainfall.getoptStorm());
                tribs_step_t s = gpu.flags();
                rains.emplace_back(H.cells.size());
                for (size_t i = 0; i < H.cells.size(); i++) rains.back()[i] = H.cells[i]->getRain();
                s.rain_mode = TRIBS_RAIN_PER_CELL;
                st.push_back(s);
                const bool gw = SimCtrl.GW_model_label && !fmod(Timer.getCurrentTime(), Timer.getGWTimeStep());
                masks.push_back(TRIBS_PHASE_UNSAT | TRIBS_PHASE_ROUTE | TRIBS_PHASE_RESET | (gw ? TRIBS_PHASE_SAT : 0));
            }
            for (size_t k = 0; k < st.size(); k++) st[k].rain = rains[k].data();
            gpu.runBatch(st, masks);
            if (gpu_channels) {
                vector<double> q = gpu.routeChannels((int)st.size());
                qout_series.insert(qout_series.end(), q.begin(), q.end());
            }
            step += (int)st.size();
            if (sw && step >= snap_from && step <= snap_to) {
                gpu.syncToNodes();
                char pre[64];
                snprintf(pre, sizeof pre, "s%d_e_", step);
                H.dump_state(*sw, pre, true);
                snap_steps.push_back(step);
            }
        }
        if (sw) { sw->i32("snap_steps", snap_steps); sw->f64("qout", qout_series); sw.reset(); }
        printf("\nref_timing {\"cells\": %ld, \"steps\": %d, \"ranks\": %d}\n", (long)H.cells.size(), step, gpu_ranks);
        return 0;
    }

    // The reference's simulation_loop body (tSimul.cpp:248-293), one call at a time.
    int step = 0;
    double t_loop = 0.0, t_unsat = 0.0, t_sat = 0.0, t_route = 0.0, t_rest = 0.0;
    vector<int32_t> snap_steps;
    while (!Timer.IsFinished() && !(max_steps > 0 && step >= max_steps)) {
        step++;
        bool snap = sw && step >= snap_from && step <= snap_to && ((step - snap_from) % snap_every == 0);
        char pre[64];
        double t0 = now_s();
        Timer.Advance(Timer.getTimeStep());
        Sim.UpdatePrecipitationInput(Rainfall.getoptStorm());
        if (gpu_et) {   // Simulator::SurfaceHydroProcesses (tSimul.cpp:419-461) with the two cell loops on the GPU
            Sim.get_next_met();
            if (Timer.getCurrentTime() == Sim.met_hour) gpu.callEvapoPotential();
            if (Timer.getCurrentTime() == Sim.eti_hour) gpu.callEvapoTrans(Intercept.getIoption() != 0);
            gpu.flushET();
            if (snap && phases.find('m') != string::npos) gpu.syncEtToNodes();
        } else if (gpu_snow) {   // the snow branch of SurfaceHydroProcesses (tSimul.cpp:464-492), cell loop on the GPU
            Sim.get_next_met();
            if (Timer.getCurrentTime() == Sim.met_hour) gpu.callSnowPack(Intercept.getIoption() != 0);
            if (snap && phases.find('m') != string::npos) gpu.syncSnowToNodes();
        } else
            Sim.SurfaceHydroProcesses(&EvapoTrans, &Intercept, &SnowPack);
        if (inject_evap > 0.0)
            for (tCNode *cn : H.cells) cn->setEvapSoil(cn->getRain() > 0.0 ? 0.0 : inject_evap);
        if (snap && phases.find('m') != string::npos) {
            snprintf(pre, sizeof pre, "s%d_m_", step); H.dump_et(*sw, pre);
        }
        double t1 = now_s();
        // SubSurfaceHydroProcesses, split so the state between the two zones is observable
        if (use_gpu) gpu.UnSaturatedZone(Timer.getTimeStep());
        else Moisture.UnSaturatedZone(Timer.getTimeStep());
        double t2 = now_s();
        if (snap && phases.find('u') != string::npos) {
            snprintf(pre, sizeof pre, "s%d_u_", step); H.dump_state(*sw, pre, false);
        }
        double t2b = now_s();
        Sim.GW_label = fmod(Timer.getCurrentTime(), Timer.getGWTimeStep());
        bool gw = false;
        if (SimCtrl.GW_model_label && !Sim.GW_label) {
            if (use_gpu) { gpu.ResetGW(); gpu.SaturatedZone(Timer.getGWTimeStep()); }
            else { Moisture.ResetGW(); Moisture.SaturatedZone(Timer.getGWTimeStep()); }
            gw = true;
        }
        double t3 = now_s();
        if (snap && gw && phases.find('g') != string::npos) {
            snprintf(pre, sizeof pre, "s%d_g_", step); H.dump_state(*sw, pre, true);
        }
        double t3b = now_s();
        if (use_gpu) gpu.syncToNodes();   // tKinemat / tFlowNet / tCOutput run on the host, unchanged
        Flow.SurfaceFlow();
        qout_series.push_back(Flow.Qout);
        double t4 = now_s();
        if (snap && phases.find('r') != string::npos) {
            snprintf(pre, sizeof pre, "s%d_r_", step); H.dump_state(*sw, pre, false);
            H.dump_routing(*sw, pre);
        }
        double t4b = now_s();
        Sim.OutputSimulatedVars(&Flow);
        Sim.UpdateWaterBalance(&Balance);
        if (snap && phases.find('b') != string::npos) {   // tWaterBalance storages, before Reset()
            snprintf(pre, sizeof pre, "s%d_b_", step);
            sw->f64(string(pre) + "UnSaturatedStorage", H.col([](tCNode *c) { return c->getUnSaturatedStorage(); }));
            sw->f64(string(pre) + "SaturatedStorage", H.col([](tCNode *c) { return c->getSaturatedStorage(); }));
            sw->f64(string(pre) + "CanopyStorVol", H.col([](tCNode *c) { return c->getCanopyStorVol(); }));
            vector<double> bs(Balance.BasinStorages, Balance.BasinStorages + 6);
            sw->f64(string(pre) + "BasinStorages", bs);
        }
        Moisture.Reset();
        if (use_gpu) gpu.Reset();
        double t5 = now_s();
        if (snap && phases.find('e') != string::npos) {
            snprintf(pre, sizeof pre, "s%d_e_", step); H.dump_state(*sw, pre, true);
        }
        if (snap) snap_steps.push_back(step);
        if (step > time_from) {
            t_rest += (t1 - t0) + (t5 - t4b);
            t_unsat += t2 - t1; t_sat += t3 - t2b; t_route += t4 - t3b;
            t_loop += (t1 - t0) + (t2 - t1) + (t3 - t2b) + (t4 - t3b) + (t5 - t4b);
        }
    }
    if (sw) { sw->i32("snap_steps", snap_steps); sw.reset(); }
    {   // outlet hydrograph of the channel router (tKinemat::Qout), one value per step
        if (gpu_et) gpu.syncEtToNodes();
        if (gpu_snow) gpu.syncSnowToNodes();
        TrbWriter qw(outdir + (use_gpu ? "/qout_gpu.trb" : "/qout_ref.trb"));
        if (EvapoTrans.getEToption() != 0) {
            qw.f64("PotEvap", H.col([](tCNode *c) { return c->getPotEvap(); }));
            qw.f64("SurfTemp", H.col([](tCNode *c) { return c->getSurfTemp(); }));
            qw.f64("CanStorage", H.col([](tCNode *c) { return c->getCanStorage(); }));
            qw.f64("CumTotEvap", H.col([](tCNode *c) { return c->getCumTotEvap(); }));
            qw.f64("CumIntercept", H.col([](tCNode *c) { return c->getCumIntercept(); }));
        }
        if (SnowPack.getSnowOpt() != 0) {
            qw.f64("IceWE", H.col([](tCNode *c) { return c->getIceWE(); }));
            qw.f64("LiqWE", H.col([](tCNode *c) { return c->getLiqWE(); }));
            qw.f64("IntSWE", H.col([](tCNode *c) { return c->getIntSWE(); }));
            qw.f64("SnTempC", H.col([](tCNode *c) { return c->getSnTempC(); }));
            qw.f64("CumMelt", H.col([](tCNode *c) { return c->getCumMelt(); }));
            qw.f64("CumIntSub", H.col([](tCNode *c) { return c->getCumIntSub(); }));
            qw.f64("PeakSWE", H.col([](tCNode *c) { return c->getPeakSWE(); }));
        }
        qw.f64("qout", qout_series);
        qw.f64("NwtNew", H.col([](tCNode *c) { return c->getNwtNew(); }));
        qw.f64("MuNew", H.col([](tCNode *c) { return c->getMuNew(); }));
        qw.f64("NfNew", H.col([](tCNode *c) { return c->getNfNew(); }));
        qw.f64("cumsrf", H.col([](tCNode *c) { return c->getCumSrf(); }));
    }
    Sim.end_simulation(&Flow);

    long ncell = (long)H.cells.size();
    printf("\nref_timing {\"cells\": %ld, \"steps\": %d, \"setup_s\": %.6f, \"loop_s\": %.6f, "
           "\"unsat_s\": %.6f, \"sat_s\": %.6f, \"route_s\": %.6f, \"other_s\": %.6f, "
           "\"cell_steps_per_s\": %.6e}\n",
           ncell, step - time_from, t_setup1 - t_setup0, t_loop, t_unsat, t_sat, t_route, t_rest,
           t_loop > 0 ? (double)ncell * (step - time_from) / t_loop : 0.0);
    fflush(stdout);
    return 0;
}
