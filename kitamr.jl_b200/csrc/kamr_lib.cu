// kamr_lib.cu — C-ABI of libkamr (include/kamr.h): context, re-flatten (topology build), state
// transfer, kernel launch sequences and the NCCL halo.
//
// Reference seam: slope!/flux!/iterate!(p4est, ka) (Flux/Slope.jl:1047, Flux/Flux.jl:458,
// Theory/Iterate.jl:5), re-flatten trigger amr_recover! (Solver/AMR.jl:54).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kamr.h"
#include "kamr_comm.h"
#include "kamr_kernels.cuh"

using namespace kamr;

#ifndef KAMR_NT
#define KAMR_NT 256
#endif
#ifndef KAMR_NT_GEN
#define KAMR_NT_GEN KAMR_NT
#endif
constexpr int NT_SLOPE = KAMR_NT_GEN;   // threads per CTA of the general slope kernel

namespace {

std::string g_create_err;

struct Fail : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            throw Fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +         \
                       std::to_string(__LINE__) + ")");                                                 \
    } while (0)
#define NCK(call)                                                                                       \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != 0) throw Fail(std::string(#call) + ": " + nccl().GetErrorString(r_));                 \
    } while (0)

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
};

#ifndef KAMR_SMALL_N
#define KAMR_SMALL_N 1024
#endif
#ifndef KAMR_CHUNK_BYTES
#define KAMR_CHUNK_BYTES 65536
#endif
constexpr int SMALL_N = KAMR_SMALL_N;   // cells up to this many points: one 128-thread CTA
constexpr int CHUNK_BYTES = KAMR_CHUNK_BYTES;   // larger cells: 256-thread CTAs whose staged f + flux planes take at most
                                                // this much shared memory (2D2F: 2048 points, 3D1F: 4096), 1..8 per cell

// One launch of a phase kernel: cells of one kernel class whose point ranges per CTA fall in one size class.
struct Bin {
    std::vector<int> cells;
    int* d_cells = nullptr;
    int C = 1;             // CTAs per cell (thread-block cluster size): 1, 2, 4 or 8
    int P = 0;             // largest point range of one CTA in this bin (sizes the staging area)
    bool wide = false;     // 256-thread CTAs (cells of more than SMALL_N points), else 128-thread CTAs
    bool stage = true;     // convected f staged in shared memory (false: in the output array; only giant cells)
    bool regular = false;  // all cells are CELL_REGULAR / CELL_REGULAR_MAPPED: phase_regular_kernel
    bool mapped = false;   // ... CELL_REGULAR_MAPPED: the MAPPED instantiation
    bool halo = false;     // cells that read ghost data: launched after the halo has arrived
    int pf_dist = 0;       // L2 prefetch distance in cells of this launch
    RegCell* d_recs = nullptr;  // regular bins: one record per cell, in launch order
};

enum KernelId { KID_SLOPE = 0, KID_MACRO_SLOPE, KID_FLUX, KID_UPDATE, KID_STEP, KID_RESIDUAL, KID_PACK, KID_UNPACK,
                KID_LIMIT, KID_SOLID_CELL, KID_SOLID_NBR, KID_STEP_REGULAR, KID_SLOPE_REGULAR, KID_STEP_REGMAP,
                KID_SLOPE_REGMAP, KID_PS_SENSOR, KID_COUNT };
const char* const kKernelNames[KID_COUNT] = {"slope_kernel", "macro_slope_kernel", "phase_kernel<FLUX>",
                                             "phase_kernel<UPDATE>", "phase_kernel<FUSED>", "residual_reduce_kernel",
                                             "pack_kernel", "unpack_kernel", "limit_kernel", "solid_cell_kernel",
                                             "solid_neighbor_kernel", "phase_regular_kernel",
                                             "slope_regular_kernel", "phase_regular_kernel<MAPPED>",
                                             "slope_regular_kernel<MAPPED>", "adaptation criteria (ps / vs)"};

struct PeerPlan {
    int rank;
    // df exchange
    std::vector<CopySeg> df_send;
    CopySeg* d_df_send = nullptr;
    long long df_send_len = 0;
    long long df_recv_off = 0, df_recv_len = 0;  // contiguous ghost range in the df array (doubles)
    // sdf exchange per level
    struct Lvl {
        std::vector<CopySeg> send, recv;
        CopySeg *d_send = nullptr, *d_recv = nullptr;
        long long send_len = 0, recv_len = 0;
    };
    std::map<int, Lvl> sdf;
    Lvl sdf_final;                            // all levels this pair does not exchange early, in one message
    unsigned long long early = 0;             // waves whose slopes this pair exchanges right after the wave
    Lvl solid;                                // df of solid ghost cells, mid-flux (Boundary/Parallel.jl:138-259)
    Lvl sw;                                   // macro slopes of the mirrors (sw_exchange!, Parallel/Ghost.jl:867)
    long long send_base = 0, recv_base = 0;  // offsets of this peer's region in the staging buffers
    int mirror_first = 0, mirror_count = 0;  // this peer's mirrors in kamr_ctx::d_send_cells
    int ghost_first = 0, ghost_count = 0;    // its ghost cells (contiguous cell ids)
};

// One-sided halo over NVLink.  Message kinds; a peer raises flag [sender rank][kind] in the receiver's
// flag table when its puts of that kind have landed.
enum { HK_DF = 0, HK_SOLID = 1, HK_SDF_FINAL = 2, HK_SW = 3, HK_SDF_EARLY = 4, HK_SLOTS = 4 + 64 };
struct HaloMsg {
    std::vector<PutSeg> segs;      // what this rank stores where
    PutSeg* d_segs = nullptr;
    std::vector<int> to;           // peer indices that receive this kind from this rank
    int* d_to = nullptr;
    std::vector<int> from;         // flag slots of the peers this rank receives this kind from
    int* d_from = nullptr;
    int epoch = 0;                 // messages of this kind sent so far (all ranks count alike)
};
struct P2P {
    bool on = false;
    int* d_flags = nullptr;                  // [nranks][HK_SLOTS], written by the peers
    std::vector<void*> opened;               // cudaIpcOpenMemHandle results (closed at the next re-flatten)
    double** d_pdf[2] = {nullptr, nullptr};  // per peer: its two df allocations
    double** d_psdf = nullptr;               // per peer: its raw-slope array
    double** d_psw = nullptr;                // per peer: its macro-slope array
    int** d_pflags = nullptr;                // per peer: its flag table
    std::map<int, HaloMsg> msg;
};

}  // namespace

struct kamr_ctx {
    kamr_config cfg{};
    GasPar gas{};
    std::string err;
    int D = 0, K = 0, M = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side_stream = nullptr;   // wall kernels of a fused step run here, beside the regular cells' phase kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // topology (host copies)
    int n_local = 0, n_ghost = 0, n_sn = 0, n_cell = 0, n_grid = 0;
    std::vector<CellInfo> cells;
    std::vector<Slot> slots;
    std::vector<FaceRec> hot;
    std::vector<int> rare;
    std::vector<long long> host_off;  // unpadded point offset per cell (host layout)
    std::vector<int> grid_n, grid_np;
    std::vector<int> grid_canon;      // first grid with identical contents: pair maps and "same grid" tests go by this id
    std::vector<long long> grid_goff, grid_hoff;
    std::vector<int8_t> h_level;      // host copy of v_level (host layout) for pair maps
    const double* h_vmid = nullptr;   // the host's v_mid, valid during kamr_upload_topology only
    std::map<std::pair<int, int>, int> rel_id;
    std::vector<long long> rel_off;
    std::vector<int> pm_start;
    // slope stages, ascending.  One rank: a single stage (regular stencils in one launch, everything else in a second
    // launch that sweeps the dependency waves with flags).  With peers: one stage per refinement level, each followed
    // by its halo exchange.
    struct SlopeStage {
        int wave = 0;
        std::vector<SlopeReg> reg;
        std::vector<SlopeRegMap> regm;   // regular stencils with neighbours on other velocity grids
        std::vector<SlopeTask> gen;
        SlopeReg* d_reg = nullptr;
        SlopeRegMap* d_regm = nullptr;
        SlopeTask* d_gen = nullptr;
        bool flags = false;   // gen tasks carry dependency lists
        int pf_reg = 0, pf_regm = 0;   // L2 prefetch distances (tasks) of the two regular launches
        int reg_int = 0, regm_int = 0, gen_int = 0;   // leading tasks of each list that read no ghost data ("interior")
    };
    std::vector<SlopeStage> slope_stages;
    std::vector<int> slope_deps;
    int slope_epoch = 0;
    unsigned* d_slope_ticket = nullptr;   // device-wide ticket counter of the slope dependency sweep
    unsigned slope_ticket_base = 0;       // its value before the next launch
    std::vector<SlopeNbr> slope_nb;
    int max_smem_optin = 0;
    bool keep_sdf = false;        // KAMR_OPT_KEEP_SDF: fused steps also write the raw slopes of every cell
    bool raw_sdf_valid = false;   // g.sdf holds the reference's sdf for every local cell
    bool sw_valid = false;        // g.sw holds the macro slopes of the current state (local cells and ghosts)
    std::vector<int> limit_cells; // local fluid + ghost fluid cells (limit_kernel after upload_aux)
    int* d_limit_cells = nullptr;
    struct MergedSegs { CopySeg *d_send = nullptr, *d_recv = nullptr; int n_send = 0, n_recv = 0; };
    std::map<std::pair<int, int>, MergedSegs> merged_segs;   // (what, level) -> pack / unpack lists of ALL peers
    std::vector<unsigned long long> peer_early;   // per peer: waves with an early slope exchange (see build_topology)
    unsigned long long early_mask = 0;            // union over the peers
    int* d_ghost_fluid = nullptr;                 // all fluid ghost cells (limited after the final slope exchange)
    int n_ghost_fluid = 0;
    std::vector<int> fluid_cells;
    int* d_fluid_cells = nullptr;
    // immersed boundary
    std::vector<SolidTask> solid_tasks;
    std::vector<SnTask> sn_tasks;
    std::vector<IbNbr> ib_nb;
    SolidTask* d_solid_tasks = nullptr;
    SnTask* d_sn_tasks = nullptr;
    std::vector<Bin> bins;
    bool padded = false;
    long long npts_pad = 0, npts_host = 0;
    // device
    std::vector<void*> allocs;
    long long device_bytes = 0;
    DevView dv{};
    double* d_res = nullptr;
    double* h_res = nullptr;  // pinned
    int* h_err = nullptr;     // pinned copy of dv.err_flag
    double* d_stage = nullptr;  // device staging for padded transfers (host layout, contiguous)
    size_t d_stage_doubles = 0;
    long long* d_host_off = nullptr;
    // halo
    P2P p2p;
    int df_parity = 0;             // which of the two df allocations is current (peers swap in lockstep)
    bool df_wait_pending = false;  // the df halo of the last step was sent; its arrival has not been waited for yet
    cudaStream_t comm_stream = nullptr;   // puts run here, beside the kernels of the main stream
    cudaEvent_t ev_put_ready = nullptr, ev_put_done = nullptr;
    bool put_in_flight = false;
    ncclComm_t comm = nullptr;
    std::vector<PeerPlan> peers;
    double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
    long long halo_bytes_step = 0;
    // physical-space adaptation sensor (kamr_ps_criterion): the neighbour lists of the host mesh, per-cell work arrays
    int *d_nb_state = nullptr, *d_nb_off = nullptr, *d_nb_ids = nullptr, *d_send_cells = nullptr;
    double *d_ps_lohner = nullptr, *d_ps_above = nullptr, *d_ps_sensor = nullptr;
    // velocity-space adaptation inputs (kamr_vs_criterion): face-neighbour tables per distinct velocity grid
    struct VsCache {
        bool valid = false;
        int maxlevel = 0, trees[3] = {0, 0, 0};
        double vmin[3] = {0, 0, 0}, vmax[3] = {0, 0, 0};
        int* d_nbt = nullptr;
        long long* d_nb_off = nullptr;
    } vs_cache;
    // partition migration device to device (kamr_migrate_begin / _finish): the arrived cells wait here across the
    // re-flatten (plain cudaMalloc, not part of the topology's allocations)
    struct Migrate {
        bool pending = false;
        double *d_df = nullptr, *d_w = nullptr, *d_prim = nullptr;
        double* d_sdf = nullptr;            // raw slopes of the cells that stay on this rank (see migrate_begin)
        long long points = 0;
        int cells = 0;
        int kept_first = 0, kept_cells = 0; // their range among the arrivals
        long long kept_points = 0;
        void release() {
            cudaFree(d_df); cudaFree(d_w); cudaFree(d_prim); cudaFree(d_sdf);
            d_df = d_w = d_prim = d_sdf = nullptr; pending = false; points = 0; cells = 0;
            kept_first = kept_cells = 0; kept_points = 0;
        }
    } mig;
    unsigned char* d_vs_flags = nullptr;   // refine_flag | coarsen_ok of the local points (host order)
    double* d_vs_res = nullptr;
    // per-kernel timing (kamr_profile_enable)
    bool profiling = false;
    struct ProfRec { int kid; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    // stats
    long long launches = 0;
    long long n_phase_local = 0;
    int fused_cells = 0;

    template <class T>
    T* dalloc(size_t n) {
        void* p = nullptr;
        size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
        CK(cudaMalloc(&p, bytes));
        allocs.push_back(p);
        device_bytes += (long long)bytes;
        return (T*)p;
    }
    template <class T>
    T* dupload(const std::vector<T>& v) {
        T* p = dalloc<T>(v.size());
        if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        return p;
    }
    void free_topology() {
        if (stream) cudaStreamSynchronize(stream);
        if (comm_stream) cudaStreamSynchronize(comm_stream);
        for (void* p : allocs) cudaFree(p);
        allocs.clear();
        device_bytes = 0;
        cells.clear(); slots.clear(); hot.clear(); rare.clear(); host_off.clear(); grid_n.clear(); grid_np.clear(); grid_goff.clear();
        grid_hoff.clear(); grid_canon.clear(); h_level.clear(); rel_id.clear(); rel_off.clear(); pm_start.clear();
        slope_stages.clear(); slope_deps.clear(); slope_nb.clear(); fluid_cells.clear(); bins.clear(); peers.clear();
        dv = DevView{};
        d_host_off = nullptr;
        d_slope_ticket = nullptr;
        d_res = nullptr; d_sendbuf = d_recvbuf = nullptr; d_fluid_cells = nullptr; d_limit_cells = nullptr;
        d_nb_state = d_nb_off = d_nb_ids = d_send_cells = nullptr;
        d_ps_lohner = d_ps_above = d_ps_sensor = nullptr;
        vs_cache = VsCache{}; d_vs_flags = nullptr; d_vs_res = nullptr;
        limit_cells.clear(); raw_sdf_valid = false; sw_valid = false;
        peer_early.clear(); early_mask = 0; d_ghost_fluid = nullptr; n_ghost_fluid = 0; merged_segs.clear();
        for (void* q : p2p.opened) cudaIpcCloseMemHandle(q);
        p2p = P2P{};
        df_parity = 0; df_wait_pending = false; put_in_flight = false;
        solid_tasks.clear(); sn_tasks.clear(); ib_nb.clear(); d_solid_tasks = nullptr; d_sn_tasks = nullptr;
    }
};

namespace {

// Counts a kernel launch and, while profiling is on, brackets it with events on the stream.
struct Launch {
    kamr_ctx* c;
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    Launch(kamr_ctx* c_, int kid, cudaStream_t st_ = nullptr) : c(c_), st(st_ ? st_ : c_->stream) {
        c->launches++;
        if (!c->profiling) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else CK(cudaEventCreate(&e));
            return e;
        };
        cudaEvent_t a = get();
        b = get();
        CK(cudaEventRecord(a, st));
        c->prof_recs.push_back({kid, a, b});
    }
    ~Launch() { if (b) cudaEventRecord(b, st); }
};

// ------------------------------------------------------------------------------------------------
// pair map of grid a onto grid b: start[i] = first point of b matched with point i of a
// in the reference's merge-walk (Flux/Slope.jl:29-64).  Integer volume accounting instead of the
// reference's floating-point `flag` accumulation; tests check it against the oracle's restatement.
int build_rel(kamr_ctx* c, int ga, int gb) {
    ga = c->grid_canon[ga]; gb = c->grid_canon[gb];   // identical grids stored twice (per-cell grids of a host that
    if (ga == gb) return -1;                          // does not deduplicate) are the same grid
    auto key = std::make_pair(ga, gb);
    auto it = c->rel_id.find(key);
    if (it != c->rel_id.end()) return it->second;
    const int D = c->D;
    const int na = c->grid_n[ga], nb = c->grid_n[gb];
    const int8_t* la = c->h_level.data() + c->grid_hoff[ga];
    const int8_t* lb = c->h_level.data() + c->grid_hoff[gb];
    int lmax = 0;
    for (int i = 0; i < na; ++i) lmax = std::max<int>(lmax, la[i]);
    for (int j = 0; j < nb; ++j) lmax = std::max<int>(lmax, lb[j]);
    auto vol = [&](int l) { return (long long)1 << (D * (lmax - l)); };
    const long long off = (long long)c->pm_start.size();
    c->pm_start.resize(off + na + 1);
    int* start = c->pm_start.data() + off;
    int j = 0;
    long long acc = 0;
    for (int i = 0; i < na; ++i) {
        if (j >= nb) throw Fail("velocity grids " + std::to_string(ga) + "/" + std::to_string(gb) +
                                " do not cover the same domain");
        start[i] = j;
        const long long va = vol(la[i]), vb = vol(lb[j]);
        if (va == vb && acc == 0) {
            ++j;
        } else if (va > vb) {
            long long got = 0;
            while (got < va) {
                if (j >= nb) throw Fail("velocity grid walk overran (coarse side)");
                got += vol(lb[j]);
                ++j;
            }
            if (got != va) throw Fail("velocity grids are not nested");
        } else {
            acc += va;
            if (acc == vb) { ++j; acc = 0; }
            else if (acc > vb) throw Fail("velocity grids are not nested");
        }
    }
    if (j != nb || acc != 0) throw Fail("velocity grid walk did not consume the neighbour grid");
    start[na] = nb;
    // Upwinding splits a face's points by the sign of v_d ON EACH SIDE'S OWN GRID (make_mask, Flux/Flux.jl); the gather
    // decides by the receiving point's sign alone, which is the same thing as long as a point and the points it is
    // matched with lie on the same side of v_d = 0, i.e. as long as v = 0 is a root-grid corner.  The reference
    // enforces exactly that in check_vs_setting (Solver/Types.jl:335-353); enforce it here too.
    if (c->h_vmid) {
        const double* va = c->h_vmid + c->grid_hoff[ga] * D;
        const double* vb = c->h_vmid + c->grid_hoff[gb] * D;
        for (int i = 0; i < na; ++i)
            for (int jj = start[i]; jj < std::max(start[i] + 1, start[i + 1]); ++jj)
                for (int d = 0; d < D; ++d)
                    if ((va[(size_t)d * na + i] > 0.) != (vb[(size_t)d * nb + jj] > 0.))
                        throw Fail("velocity grids " + std::to_string(ga) + "/" + std::to_string(gb) +
                                   ": matched points lie on different sides of v = 0; the origin must be a root-grid "
                                   "corner of the velocity space (check_vs_setting, Solver/Types.jl:335)");
    }
    const int id = (int)c->rel_off.size();
    c->rel_off.push_back(off);
    c->rel_id[key] = id;
    return id;
}

double face_area_of(const CellInfo& ci, int D, int dir) {  // Flux/Flux.jl:10-15
    if (D == 2) return ci.ds[dir == 0 ? 1 : 0];
    if (dir == 0) return ci.ds[1] * ci.ds[2];
    if (dir == 1) return ci.ds[0] * ci.ds[2];
    return ci.ds[0] * ci.ds[1];
}

struct NbList {
    int cnt;
    const int32_t* ids;
};

// _ps_has_transverse_offset, Flux/Slope.jl:177-201
bool has_offset(const kamr_ctx* c, int cell, NbList nb, int dir) {
    const int D = c->D;
    if (nb.cnt == 0) return false;
    for (int t = 0; t < D; ++t) {
        if (t == dir) continue;
        double avg = 0.0;
        for (int j = 0; j < nb.cnt; ++j) avg += c->cells[nb.ids[j]].mid[t];
        avg /= nb.cnt;
        if (avg != c->cells[cell].mid[t]) return true;
    }
    return false;
}

long long rel_offset(kamr_ctx* c, int ga, int gb) {
    const int id = build_rel(c, ga, gb);
    return id < 0 ? -1 : c->rel_off[id];
}

// appends the neighbours of one side to c->slope_nb; returns the count
int fill_side(kamr_ctx* c, int cell, NbList nb, int dir, int project /*0 no,1 yes*/) {
    if (nb.cnt > 4) throw Fail("more than 4 neighbours on one face");
    for (int a = 0; a < nb.cnt; ++a) {
        const int nbr = nb.ids[a];
        const CellInfo& cn = c->cells[nbr];
        SlopeNbr e;
        memset(&e, 0, sizeof(e));
        e.doff = cn.doff; e.goff = cn.goff; e.np = cn.np;
        e.rel_off = rel_offset(c, c->cells[cell].grid, cn.grid);
        if (project) {
            bool any = false;
            for (int t = 0; t < c->D; ++t) {
                e.dm[t] = (t == dir) ? 0.0 : (c->cells[cell].mid[t] - cn.mid[t]);
                any = any || e.dm[t] != 0.0;
            }
            e.proj = any ? 1 : 0;  // dm == 0 contributes exactly 0 in the reference as well
        }
        c->slope_nb.push_back(e);
    }
    return nb.cnt;
}

void set_bound(kamr_ctx* c, int cell, NbList nb, double ds, int dir, int project, SlopeDir& out) {
    out.mode = SLOPE_BOUND;
    out.nb_begin = (int)c->slope_nb.size();
    out.nA = fill_side(c, cell, nb, dir, project);
    out.nB = 0;
    out.invA = 1.0 / (ds * out.nA);
    out.invB = 0.0;
}
void set_inner(kamr_ctx* c, int cell, NbList L, NbList R, double dsL, double dsR, int dir, int projL, int projR,
               SlopeDir& out) {
    out.mode = SLOPE_INNER;
    out.nb_begin = (int)c->slope_nb.size();
    out.nA = fill_side(c, cell, L, dir, projL);
    out.nB = fill_side(c, cell, R, dir, projR);
    out.invA = 1.0 / (dsL * out.nA);
    out.invB = 1.0 / (dsR * out.nB);
}

// the 15 update_slope! methods (Flux/Slope.jl:458-771) resolved into a stencil descriptor
void baseline_dir(kamr_ctx* c, const kamr_mesh* m, int cell, int dir, SlopeDir& out) {
    const int D = c->D;
    const int e = cell * 2 * D + 2 * dir;
    const int sL = m->nb_state[e], sR = m->nb_state[e + 1];
    NbList L{m->nb_off[e + 1] - m->nb_off[e], m->nb_ids + m->nb_off[e]};
    NbList R{m->nb_off[e + 2] - m->nb_off[e + 1], m->nb_ids + m->nb_off[e + 1]};
    const CellInfo& ci = c->cells[cell];
    auto midd = [&](int id) { return c->cells[id].mid[dir]; };
    memset(&out, 0, sizeof(out));
    if (sL == 1 && sR == 1) {
        const bool solidL = c->cells[L.ids[0]].bound_enc < 0, solidR = c->cells[R.ids[0]].bound_enc < 0;
        if (solidL && solidR) { out.mode = SLOPE_ZERO; return; }
        if (solidL) { set_bound(c, cell, R, ci.mid[dir] - midd(R.ids[0]), dir, 0, out); return; }
        if (solidR) { set_bound(c, cell, L, ci.mid[dir] - midd(L.ids[0]), dir, 0, out); return; }
        set_inner(c, cell, L, R, ci.ds[dir], -ci.ds[dir], dir, 0, 0, out);
        return;
    }
    if (sL == 0 && sR == 0) throw Fail("cell with domain boundaries on both sides of one direction (no such "
                                       "update_slope! method in the reference)");
    if (sL == 0) { set_bound(c, cell, R, ci.mid[dir] - midd(R.ids[0]), dir, 0, out); return; }
    if (sR == 0) { set_bound(c, cell, L, ci.mid[dir] - midd(L.ids[0]), dir, 0, out); return; }
    const double ds = ci.ds[dir];
    const double dsL = (sL == 1) ? ds : ((sL == -1 ? 1.5 : 0.75) * ds);
    const double dsR = (sR == 1) ? -ds : (-(sR == -1 ? 1.5 : 0.75) * ds);
    set_inner(c, cell, L, R, dsL, dsR, dir, 0, 0, out);
}

// update_slope_transverse_level!, Flux/Slope.jl:849-945
void transverse_dir(kamr_ctx* c, const kamr_mesh* m, int cell, int dir, SlopeDir& out) {
    const int D = c->D;
    const int e = cell * 2 * D + 2 * dir;
    const int sL = m->nb_state[e], sR = m->nb_state[e + 1];
    if (!(sL == -1 || sR == -1)) { baseline_dir(c, m, cell, dir, out); return; }
    NbList L{m->nb_off[e + 1] - m->nb_off[e], m->nb_ids + m->nb_off[e]};
    NbList R{m->nb_off[e + 2] - m->nb_off[e + 1], m->nb_ids + m->nb_off[e + 1]};
    const CellInfo& ci = c->cells[cell];
    memset(&out, 0, sizeof(out));
    auto solid = [&](NbList l) { return c->cells[l.ids[0]].bound_enc < 0; };
    if (sL != 0 && sR != 0) {
        const double dsL = ci.mid[dir] - c->cells[L.ids[0]].mid[dir];
        const double dsR = ci.mid[dir] - c->cells[R.ids[0]].mid[dir];
        if (solid(L)) { set_bound(c, cell, R, dsR, dir, 1, out); return; }
        if (solid(R)) { set_bound(c, cell, L, dsL, dir, 1, out); return; }
        set_inner(c, cell, L, R, dsL, dsR, dir, has_offset(c, cell, L, dir) ? 1 : 0,
                  has_offset(c, cell, R, dir) ? 1 : 0, out);
    } else if (sR == 0 && sL == -1) {
        if (solid(L)) { out.mode = SLOPE_KEEP; return; }
        set_bound(c, cell, L, ci.mid[dir] - c->cells[L.ids[0]].mid[dir], dir, 1, out);
    } else {
        if (solid(R)) { out.mode = SLOPE_KEEP; return; }
        set_bound(c, cell, R, ci.mid[dir] - c->cells[R.ids[0]].mid[dir], dir, 1, out);
    }
}

void setup_p2p(kamr_ctx* c, const kamr_mesh* m);

void build_topology(kamr_ctx* c, const kamr_mesh* m) {
    const int D = c->D, K = c->K, M = c->M;
    const bool verbose = getenv("KAMR_VERBOSE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {   // KAMR_VERBOSE: where a re-flatten spends its time
        if (!verbose) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[kamr] re-flatten %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    c->free_topology();
    lap("free old topology");
    c->n_local = m->n_local; c->n_ghost = m->n_ghost; c->n_sn = m->n_solidnbr;
    c->n_cell = m->n_local + m->n_ghost + m->n_solidnbr;
    c->n_grid = m->n_grid;
    if (c->n_local <= 0) throw Fail("mesh has no local cells");
    if (m->n_ghost < 0 || m->n_solidnbr < 0 || m->n_peer < 0 || m->n_face < 0 || m->n_grid <= 0)
        throw Fail("negative count in kamr_mesh");
    // halo and immersed-boundary index arrays index host tables below: range-check them before first use
    if (m->n_peer > 0) {
        // the early-level handshake below makes this call collective over the pair communicator: every rank must have
        // called kamr_comm_init first, or the two sides would build different exchange schedules (or block)
        if (c->cfg.nranks > 1 && !c->comm)
            throw Fail("mesh has peers: call kamr_comm_init on every rank before kamr_upload_topology");
        if (m->send_off[0] != 0 || m->recv_off[0] != 0) throw Fail("send_off / recv_off must start at 0");
        for (int p = 0; p < m->n_peer; ++p) {
            if (m->peer_rank[p] < 0 || m->peer_rank[p] >= std::max(1, c->cfg.nranks) || m->peer_rank[p] == c->cfg.rank)
                throw Fail("peer_rank out of range");
            if (m->send_off[p + 1] < m->send_off[p] || m->recv_off[p + 1] < m->recv_off[p])
                throw Fail("send_off / recv_off must ascend");
        }
        if (m->recv_off[m->n_peer] > m->n_ghost) throw Fail("recv_off exceeds n_ghost");
        for (int q = 0; q < m->send_off[m->n_peer]; ++q)
            if (m->send_cells[q] < 0 || m->send_cells[q] >= m->n_local) throw Fail("send_cells must be local cell ids");
    }
    if (m->ib) {
        const kamr_ib* ib = m->ib;
        if (ib->n_solid < 0 || ib->n_sn < 0) throw Fail("negative count in kamr_ib");
        for (int q = 0; q < (ib->n_solid ? ib->solid_nb_off[ib->n_solid] : 0); ++q)
            if (ib->solid_nb_ids[q] < 0 || ib->solid_nb_ids[q] >= c->n_cell) throw Fail("solid_nb_ids out of range");
        for (int q = 0; q < (ib->n_sn ? ib->sn_nb_off[ib->n_sn] : 0); ++q)
            if (ib->sn_nb_ids[q] < 0 || ib->sn_nb_ids[q] >= c->n_cell) throw Fail("sn_nb_ids out of range");
        for (int q = 0; q < ib->n_sn; ++q)
            if (ib->sn_donor[q] < 0 || ib->sn_donor[q] >= m->n_local) throw Fail("sn_donor must be a local cell");
    }
    // ---- grids
    c->grid_n.resize(m->n_grid); c->grid_np.resize(m->n_grid);
    c->grid_goff.resize(m->n_grid + 1); c->grid_hoff.resize(m->n_grid + 1);
    c->grid_goff[0] = 0; c->grid_hoff[0] = 0;
    bool padded = false;
    for (int g = 0; g < m->n_grid; ++g) {
        const long long n = m->grid_off[g + 1] - m->grid_off[g];
        if (n <= 0 || n > (1ll << 30)) throw Fail("bad velocity grid size");
        c->grid_n[g] = (int)n;
        c->grid_np[g] = (int)round_up(n, PAD);
        padded = padded || (c->grid_np[g] != n);
        c->grid_hoff[g + 1] = c->grid_hoff[g] + n;
        c->grid_goff[g + 1] = c->grid_goff[g] + c->grid_np[g];
    }
    c->padded = padded;
    {   // canonical ids: grids with the same levels and midpoints are one grid for every index map (the host may hand
        // every cell its own copy, as VsData objects are per cell in the reference)
        c->grid_canon.resize(m->n_grid);
        std::map<std::pair<int, unsigned long long>, std::vector<int>> buckets;
        for (int g = 0; g < m->n_grid; ++g) {
            const int n = c->grid_n[g];
            const int8_t* lv = m->v_level + c->grid_hoff[g];
            const double* vm = m->v_mid + c->grid_hoff[g] * D;
            unsigned long long h = 1469598103934665603ull;
            for (int i = 0; i < n; ++i) h = (h ^ (unsigned char)lv[i]) * 1099511628211ull;
            unsigned long long b0, b1;
            memcpy(&b0, vm, 8); memcpy(&b1, vm + (size_t)D * n - 1, 8);
            h = (h ^ b0) * 1099511628211ull; h = (h ^ b1) * 1099511628211ull;
            auto& bk = buckets[std::make_pair(n, h)];
            int canon = g;
            for (int g2 : bk) {
                if (memcmp(lv, m->v_level + c->grid_hoff[g2], n) == 0 &&
                    memcmp(vm, m->v_mid + c->grid_hoff[g2] * D, sizeof(double) * D * n) == 0) { canon = g2; break; }
            }
            if (canon == g) bk.push_back(g);
            c->grid_canon[g] = canon;
        }
    }
    const long long gpts_h = c->grid_hoff[m->n_grid], gpts_d = c->grid_goff[m->n_grid];
    c->h_level.assign(m->v_level, m->v_level + gpts_h);
    c->h_vmid = m->v_mid;
    {
        std::vector<int8_t> lv(gpts_d, 0);
        std::vector<unsigned char> sg(gpts_d, 0);
        std::vector<double> wt(gpts_d, 0.0), vm((size_t)gpts_d * D, 0.0);
        for (int g = 0; g < m->n_grid; ++g) {
            const int n = c->grid_n[g], np = c->grid_np[g];
            const long long ho = c->grid_hoff[g], go = c->grid_goff[g];
            memcpy(lv.data() + go, m->v_level + ho, n);
            memcpy(wt.data() + go, m->v_weight + ho, sizeof(double) * n);
            for (int d = 0; d < D; ++d)
                memcpy(vm.data() + go * D + (size_t)d * np, m->v_mid + ho * D + (size_t)d * n, sizeof(double) * n);
            // padding points replicate the last real point so padded lanes stay finite
            for (int i = n; i < np; ++i) {
                lv[go + i] = lv[go + n - 1];
                for (int d = 0; d < D; ++d) vm[go * D + (size_t)d * np + i] = vm[go * D + (size_t)d * np + n - 1];
            }
        }
        // bit d of v_sign: v_d > 0, i.e. the neighbour across the LOW face of direction d is upwind.  One byte per
        // point and per distinct grid, so the phase kernels find their upwind records from L1 instead of waiting for v
        for (int g = 0; g < m->n_grid; ++g) {
            const int np = c->grid_np[g];
            const long long go = c->grid_goff[g];
            for (int i = 0; i < np; ++i)
                for (int d = 0; d < D; ++d)
                    if (vm[go * D + (size_t)d * np + i] > 0.) sg[go + i] |= (unsigned char)(1u << d);
        }
        {   // packed statics: per-axis tables of the distinct coordinates, weight by level, one word per point
            std::vector<std::vector<double>> tab(D);
            for (int d = 0; d < D; ++d) {
                std::vector<double>& t = tab[d];
                for (int g = 0; g < m->n_grid; ++g) {
                    if (c->grid_canon[g] != g) continue;
                    const double* p = vm.data() + c->grid_goff[g] * D + (size_t)d * c->grid_np[g];
                    t.insert(t.end(), p, p + c->grid_n[g]);
                    std::sort(t.begin(), t.end());
                    t.erase(std::unique(t.begin(), t.end()), t.end());
                }
                if ((int)t.size() > (1 << VPK_BITS))
                    throw Fail("more than " + std::to_string(1 << VPK_BITS) + " distinct velocity coordinates on one axis");
            }
            int ntab = 1;
            for (int d = 0; d < D; ++d) ntab = std::max<int>(ntab, (int)tab[d].size());
            ntab = (int)round_up(ntab, 2);
            std::vector<double> vt((size_t)D * ntab + VPK_LEVELS, 0.0);
            for (int d = 0; d < D; ++d) std::copy(tab[d].begin(), tab[d].end(), vt.begin() + (size_t)d * ntab);
            std::vector<char> have(VPK_LEVELS, 0);
            std::vector<unsigned> pk(gpts_d, 0u);
            for (int g = 0; g < m->n_grid; ++g) {
                const int np = c->grid_np[g];
                const long long go = c->grid_goff[g];
                const int cg = c->grid_canon[g];
                if (cg != g) {   // identical contents: copy the words
                    std::copy(pk.begin() + c->grid_goff[cg], pk.begin() + c->grid_goff[cg] + np, pk.begin() + go);
                    continue;
                }
                for (int i = 0; i < np; ++i) {
                    const int l = lv[go + i];
                    if (l < 0 || l >= VPK_LEVELS) throw Fail("velocity level out of range (0..15)");
                    unsigned w = (unsigned)l << 27;
                    for (int d = 0; d < D; ++d) {
                        const double x = vm[go * D + (size_t)d * np + i];
                        const int k = (int)(std::lower_bound(tab[d].begin(), tab[d].end(), x) - tab[d].begin());
                        w |= (unsigned)k << (VPK_BITS * d);
                    }
                    pk[go + i] = w;
                    if (i < c->grid_n[g]) {
                        const double wgt = wt[go + i];
                        if (!have[l]) { have[l] = 1; vt[(size_t)D * ntab + l] = wgt; }
                        else if (vt[(size_t)D * ntab + l] != wgt)
                            throw Fail("velocity weights are not a function of the level (VsData.weight = root weight / "
                                       "2^(DIM*level), Velocity_space/Rebuild.jl:60)");
                    }
                }
            }
            c->dv.v_pack = c->dupload(pk);
            c->dv.v_tab = c->dupload(vt);
            c->dv.n_vtab = ntab;
            CK(cudaStreamSynchronize(c->stream));
        }
        c->dv.v_sign = c->dupload(sg);
        c->dv.v_level = c->dupload(lv);
        c->dv.v_weight = c->dupload(wt);
        c->dv.v_mid = c->dupload(vm);
        CK(cudaStreamSynchronize(c->stream));
    }
    lap("grids + packed statics");
    // ---- cells
    c->cells.resize(c->n_cell);
    c->host_off.resize(c->n_cell + 1);
    c->host_off[0] = 0;
    long long doff = 0;
    c->n_phase_local = 0;
    for (int i = 0; i < c->n_cell; ++i) {
        CellInfo& ci = c->cells[i];
        memset(&ci, 0, sizeof(ci));
        const int g = m->cell_grid[i];
        if (g < 0 || g >= m->n_grid) throw Fail("cell_grid out of range");
        ci.grid = g; ci.n = c->grid_n[g]; ci.np = c->grid_np[g];
        ci.doff = doff; ci.goff = c->grid_goff[g];
        ci.bound_enc = m->bound_enc[i];
        ci.ps_level = m->ps_level[i];
        ci.vol = 1.0;
        for (int d = 0; d < D; ++d) {
            ci.ds[d] = m->ds[(size_t)i * D + d];
            ci.mid[d] = m->mid[(size_t)i * D + d];
            ci.vol *= ci.ds[d];  // reduce(*, ps_data.ds), Iterate.jl:106
        }
        doff += ci.np;
        c->host_off[i + 1] = c->host_off[i] + ci.n;
        if (i < c->n_local && ci.bound_enc >= 0) c->n_phase_local += ci.n;
    }
    c->npts_pad = doff; c->npts_host = c->host_off[c->n_cell];
    lap("cells");
    // ---- slots (cell-centric view of the face list)
    std::vector<std::vector<Slot>> per_cell(c->n_local);
    for (int f = 0; f < m->n_face; ++f) {
        const int kind = m->face_kind[f], here = m->face_here[f], there = m->face_there[f], dir = m->face_dir[f];
        const double rot = m->face_rot[f];
        if (here < 0 || here >= c->n_local) throw Fail("face_here must be a local cell");
        if (dir < 0 || dir >= D) throw Fail("face_dir out of range");
        Slot s;
        memset(&s, 0, sizeof(s));
        s.dir = dir; s.rot = rot; s.face = f; s.is_here = 1; s.rel_off = -1;
        const CellInfo& ch = c->cells[here];
        for (int t = 0; t < D; ++t) { s.fmid[t] = m->face_mid[(size_t)f * D + t]; s.own_mid[t] = ch.mid[t]; }
        double area = face_area_of(ch, D, dir);
        if (kind == KAMR_FACE_DOMAIN) {
            if (there < 0 || there >= m->n_bc) throw Fail("domain face bc index out of range");
            s.nbr = -1;
            s.area = rot * area;
            switch (m->bc_type[there]) {
                case KAMR_BC_MAXWELLIAN:
                    // calc_domain_flux(DVM, Maxwellian) reads undefined variables in the reference (Flux/DVM.jl:3,12)
                    if (c->gas.flux_type == KAMR_FLUX_DVM)
                        throw Fail("DVM flux with a Maxwellian domain wall cannot run in the reference (Flux/DVM.jl:12)");
                    s.kind = SLOT_BC_MAXWELL; break;
                case KAMR_BC_SUPERSONIC_INFLOW: s.kind = SLOT_BC_INFLOW; break;
                case KAMR_BC_UNIFORM_OUTFLOW: s.kind = SLOT_BC_UNIFORM; break;
                case KAMR_BC_INTERPOLATED_OUTFLOW:
                    if (D != 2) throw Fail("InterpolatedOutflow is 2-D only (CAIDVM.jl:71)");
                    s.kind = SLOT_BC_INTERP; break;
                default: throw Fail("unknown bc type");
            }
            for (int q = 0; q < M; ++q) s.bc[q] = m->bc_prim[(size_t)there * M + q];
            per_cell[here].push_back(s);
            continue;
        }
        if (there < 0 || there >= c->n_cell) throw Fail("face_there out of range");
        if (kind == KAMR_FACE_HANGING) area = area / (double)(1 << (D - 1)) * rot;  // Flux.jl:84-86
        else area = area * rot;                                                     // Flux.jl:87-89
        const CellInfo& ct = c->cells[there];
        s.nbr = there;
        s.nbr_doff = ct.doff; s.nbr_goff = ct.goff; s.nbr_np = ct.np;
        s.kind = (ct.bound_enc < 0) ? SLOT_NBR_SOLID : SLOT_INNER;
        s.rel_off = rel_offset(c, ch.grid, ct.grid);
        s.area = area;
        for (int t = 0; t < D; ++t) { s.nbr_mid[t] = m->face_there_mid[(size_t)f * D + t]; s.nds[t] = ct.ds[t]; }
        per_cell[here].push_back(s);
        if (there < c->n_local && ct.bound_enc >= 0) {  // update_macro_flux!/update_micro_flux! write-back side
            Slot r = s;
            r.is_here = 0; r.nbr = here; r.kind = SLOT_INNER;
            r.nbr_doff = ch.doff; r.nbr_goff = ch.goff; r.nbr_np = ch.np;
            r.rel_off = rel_offset(c, ct.grid, ch.grid);
            r.area = -area;
            for (int t = 0; t < D; ++t) { r.own_mid[t] = s.nbr_mid[t]; r.nbr_mid[t] = ch.mid[t]; r.nds[t] = ch.ds[t]; }
            per_cell[there].push_back(r);
        }
    }
    // side of the cell a slot's face sits on: here & rot=+1 -> low face; the there side sees the same face from across
    auto side_of = [](const Slot& s) { return (s.is_here ? (s.rot > 0) : (s.rot < 0)) ? 0 : 1; };
    for (int i = 0; i < c->n_local; ++i) {
        if ((int)per_cell[i].size() > max_slots(D)) throw Fail("too many faces on one cell");
        std::stable_sort(per_cell[i].begin(), per_cell[i].end(), [&](const Slot& a, const Slot& b) {
            return 2 * a.dir + side_of(a) < 2 * b.dir + side_of(b);
        });
        CellInfo& ci = c->cells[i];
        ci.slot_begin = (int)c->slots.size();
        c->slots.insert(c->slots.end(), per_cell[i].begin(), per_cell[i].end());
        ci.slot_end = (int)c->slots.size();
        ci.hot_begin = (int)c->hot.size();
        ci.rare_begin = (int)c->rare.size();
        int key = 0, nh = 0;
        for (int q = 0; q < (int)per_cell[i].size(); ++q) {
            const Slot& sl = per_cell[i][q];
            if (sl.kind == SLOT_BC_MAXWELL) ci.flags |= CELL_HAS_MAXWELL_WALL;
            if (sl.kind == SLOT_INNER) {
                while (key <= 2 * sl.dir + side_of(sl)) ci.side_begin[key++] = (unsigned char)nh;
                FaceRec h;
                memset(&h, 0, sizeof(h));
                h.nf_off = sl.nbr_doff * K; h.nsl_off = sl.nbr_doff * K * D; h.np = sl.nbr_np;
                h.flags = sl.rel_off < 0 ? 2 : 0;
                h.rel_off = sl.rel_off; h.ngoff = sl.nbr_goff;
                if (sl.rel_off >= 0) ci.flags |= CELL_HAS_MAPPED;
                h.area = sl.area;
                for (int t = 0; t < D; ++t) { h.fmid[t] = sl.fmid[t]; h.own_mid[t] = sl.own_mid[t]; h.nbr_mid[t] = sl.nbr_mid[t]; }
                c->hot.push_back(h);
                ++nh;
            }
            if (sl.kind != SLOT_INNER) c->rare.push_back(q);   // pair-mapped fluid faces are gathered in pass A
        }
        while (key <= 2 * D) ci.side_begin[key++] = (unsigned char)nh;
        ci.rare_count = (int)c->rare.size() - ci.rare_begin;
        bool regular = ci.rare_count == 0 && nh == 2 * D && ci.bound_enc >= 0;
        bool any_mapped = false;
        for (int k2 = 0; regular && k2 < 2 * D; ++k2) {
            const FaceRec& h = c->hot[ci.hot_begin + k2];
            regular = ci.side_begin[k2] == k2;
            any_mapped = any_mapped || !(h.flags & 2);
            if ((h.flags & 2) && h.np != ci.np) regular = false;
            // RegCell carries the normal coordinate of a side only: everything else must equal the cell's midpoint
            for (int t = 0; regular && t < D; ++t) {
                regular = h.own_mid[t] == ci.mid[t];
                if (t != k2 / 2) regular = regular && h.fmid[t] == ci.mid[t] && h.nbr_mid[t] == ci.mid[t];
            }
        }
        if (regular) ci.flags |= any_mapped ? CELL_REGULAR_MAPPED : CELL_REGULAR;
    }
    lap("slots");
    // ---- slope tasks in dependency waves (slope!, Slope.jl:1047-1070).  The reference sweeps the levels
    // coarse to fine because a fine cell next to a coarse one projects the coarse cell's FINISHED slopes
    // (Slope.jl:165-175).  Only those cells depend on anything: a cell none of whose stencils projects
    // reads df only and can run in wave 0 whatever its level; a projecting cell of level L runs in wave
    // L - Lmin (its coarse neighbours are done by then).  With peers every wave is followed by its halo
    // exchange (slope_exchange_level!, Parallel/Ghost.jl:896), so there the waves are the reference's
    // levels for every cell: both sides of a partition boundary must agree on when a cell is final.
    {
        std::map<std::pair<int, int>, std::vector<SlopeTask>> by_wave;  // (wave, pair-mapped?) -> tasks
        const bool by_level = m->n_peer > 0;
        std::vector<SlopeTask> tasks;
        std::vector<char> gen, dep;
        std::vector<std::vector<int>> deps;       // cells whose finished sdf this task projects
        std::vector<char> need_raw(c->n_cell, 0);
        std::vector<int> task_of(c->n_cell, -1);
        std::map<long long, int> cell_of_doff;
        for (int i = 0; i < c->n_cell; ++i) cell_of_doff[c->cells[i].doff] = i;
        for (int i = 0; i < c->n_local; ++i) {
            if (c->cells[i].bound_enc < 0) continue;
            const int L = c->cells[i].ps_level;
            SlopeTask t;
            memset(&t, 0, sizeof(t));
            t.cell = i;
            bool g_ = false;
            std::vector<int> dl;
            for (int d = 0; d < D; ++d) {
                if (L <= m->ps_minlevel) baseline_dir(c, m, i, d, t.d[d]);
                else transverse_dir(c, m, i, d, t.d[d]);
                if (t.d[d].nA + t.d[d].nB > 2 * (1 << (D - 1))) throw Fail("slope stencil too large");
                if (t.d[d].mode == SLOPE_KEEP) need_raw[i] = 1;
                for (int a = 0; a < t.d[d].nA + t.d[d].nB; ++a) {
                    const SlopeNbr& e = c->slope_nb[t.d[d].nb_begin + a];
                    if (e.proj) {
                        // A true dependency only if the target is coarser.  (Rounding in the averaged midpoints of a
                        // finer neighbour list can also switch the projection on, with dm ~ 1 ulp; the reference
                        // then reads the finer cells' slopes of the PREVIOUS sweep, as they are processed later —
                        // reproduced here: they run in a later wave and keep their raw sdf.)
                        const int tgt = cell_of_doff[e.doff];
                        need_raw[tgt] = 1;
                        if (c->cells[tgt].ps_level < L) dl.push_back(tgt);
                    }
                    g_ = g_ || e.rel_off >= 0;
                }
            }
            for (int sidx = c->cells[i].slot_begin; sidx < c->cells[i].slot_end; ++sidx)
                if (c->slots[sidx].kind != SLOT_INNER) need_raw[i] = 1;
            task_of[i] = (int)tasks.size();
            tasks.push_back(t); gen.push_back(g_); dep.push_back(!dl.empty()); deps.push_back(std::move(dl));
        }
        for (int p = 0; p < m->n_peer; ++p)  // mirrors: their raw slopes travel (slope_exchange_level!)
            for (int q = m->send_off[p]; q < m->send_off[p + 1]; ++q) need_raw[m->send_cells[q]] = 1;
        if (m->ib) {  // the wall kernels extrapolate with the raw slopes of the fluid cells around the body
            const kamr_ib* ib = m->ib;
            for (int q = 0; q < ib->solid_nb_off[ib->n_solid]; ++q) need_raw[ib->solid_nb_ids[q]] = 1;
            for (int q = 0; q < ib->sn_nb_off[ib->n_sn]; ++q) need_raw[ib->sn_nb_ids[q]] = 1;
            for (int q = 0; q < ib->n_sn; ++q) need_raw[ib->sn_donor[q]] = 1;
        }
        // wave of a task: with peers the reference's level sweep (both sides of a partition boundary must agree
        // on when a cell is final); on one rank the true dependency depth
        std::vector<int> wave(tasks.size(), -1);
        std::function<int(int)> depth = [&](int ti) -> int {
            if (wave[ti] >= 0) return wave[ti];
            if (wave[ti] == -2) throw Fail("cyclic slope dependency");
            wave[ti] = -2;
            int w = 0;
            for (int tgt : deps[ti]) {
                const int tj = (tgt < c->n_local) ? task_of[tgt] : -1;
                if (tj >= 0) w = std::max(w, depth(tj) + 1);
            }
            return wave[ti] = w;
        };
        auto is_regular = [&](const SlopeTask& t) {
            const CellInfo& ci = c->cells[t.cell];
            for (int d = 0; d < D; ++d) {
                const SlopeDir& sd = t.d[d];
                if (sd.mode != SLOPE_INNER || sd.nA != 1 || sd.nB != 1) return false;
                for (int a = 0; a < 2; ++a) {
                    const SlopeNbr& e = c->slope_nb[sd.nb_begin + a];
                    if (e.rel_off >= 0 || e.proj || e.np != ci.np) return false;
                }
            }
            return true;
        };
        auto is_regmapped = [&](const SlopeTask& t) {   // regular stencil, neighbours possibly on other grids
            for (int d = 0; d < D; ++d) {
                const SlopeDir& sd = t.d[d];
                if (sd.mode != SLOPE_INNER || sd.nA != 1 || sd.nB != 1) return false;
                for (int a = 0; a < 2; ++a)
                    if (c->slope_nb[sd.nb_begin + a].proj) return false;
            }
            return true;
        };
        auto make_reg = [&](const SlopeTask& t) {
            const CellInfo& ci = c->cells[t.cell];
            SlopeReg r;
            memset(&r, 0, sizeof(r));
            r.doff = ci.doff; r.n = ci.n; r.np = ci.np; r.flags = t.flags;
            for (int d = 0; d < D; ++d) {
                r.ds[d] = ci.ds[d];
                r.nb_doff[2 * d] = c->slope_nb[t.d[d].nb_begin].doff;
                r.nb_doff[2 * d + 1] = c->slope_nb[t.d[d].nb_begin + 1].doff;
                r.inv[2 * d] = t.d[d].invA; r.inv[2 * d + 1] = t.d[d].invB;
            }
            return r;
        };
        std::map<int, kamr_ctx::SlopeStage> stages;
        std::vector<int> order(tasks.size());
        for (size_t ti = 0; ti < tasks.size(); ++ti) {
            order[ti] = (int)ti;
            const int L = c->cells[tasks[ti].cell].ps_level;
            wave[ti] = by_level ? std::max(0, L - m->ps_minlevel) : depth((int)ti);
            tasks[ti].flags = need_raw[tasks[ti].cell] ? 1 : 0;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return wave[a] < wave[b]; });
        // ---- with peers: which waves need their slope halo RIGHT AFTER the wave?  Only those whose mirror slopes a
        // peer projects in a later wave (a fine cell next to a coarser ghost cell, Slope.jl:165-175).  Every other
        // mirror's slopes are needed by the flux gather only and travel in one message after the last wave.  Each rank
        // knows which ghost slopes it projects; one 8-byte handshake per pair makes the schedule symmetric.
        c->peer_early.assign(m->n_peer, 0);
        c->early_mask = 0;
        if (by_level) {
            std::vector<unsigned long long> need(m->n_peer, 0), got(m->n_peer, ~0ull);
            for (size_t ti = 0; ti < tasks.size(); ++ti)
                for (int tgt : deps[ti]) {
                    if (tgt < c->n_local) continue;
                    const int gi = tgt - c->n_local;
                    for (int p = 0; p < m->n_peer; ++p)
                        if (gi >= m->recv_off[p] && gi < m->recv_off[p + 1])
                            need[p] |= 1ull << std::min(63, std::max(0, c->cells[tgt].ps_level - m->ps_minlevel));
                }
            if (c->comm) {
                unsigned long long* d_hs = c->dalloc<unsigned long long>(2 * (size_t)m->n_peer);
                CK(cudaMemcpyAsync(d_hs, need.data(), sizeof(unsigned long long) * m->n_peer, cudaMemcpyHostToDevice,
                                   c->stream));
                NCK(nccl().GroupStart());
                for (int p = 0; p < m->n_peer; ++p) {
                    NCK(nccl().Send(d_hs + p, 1, ncclFloat64, m->peer_rank[p], c->comm, c->stream));
                    NCK(nccl().Recv(d_hs + m->n_peer + p, 1, ncclFloat64, m->peer_rank[p], c->comm, c->stream));
                }
                NCK(nccl().GroupEnd());
                CK(cudaMemcpyAsync(got.data(), d_hs + m->n_peer, sizeof(unsigned long long) * m->n_peer,
                                   cudaMemcpyDeviceToHost, c->stream));
                CK(cudaStreamSynchronize(c->stream));
            }   // without a communicator yet: every wave early, as slope_exchange_level! does
            for (int p = 0; p < m->n_peer; ++p) {
                c->peer_early[p] = c->comm ? (need[p] | got[p]) : ~0ull;
                c->early_mask |= c->peer_early[p];
            }
        }
        // launches: stage k runs before the k-th early exchange.  A task belongs to the first stage in which everything it
        // projects is final: stage 0 unless it (transitively) projects the slopes of a ghost cell, which arrive with the
        // early exchange of the ghost's wave.  So nearly every task of every wave sits in stage 0 (dependencies among
        // local tasks go through the per-cell epoch flags) and the later stages hold the few cells behind a ghost; a
        // mirror of early wave e only ever projects coarser cells, whose exchanges come before e's, so it is final in time.
        auto exch_index = [&](int w) {   // position of wave w among the early waves
            int k = 0;
            for (int e = 0; e < w && e < 64; ++e)
                if (c->early_mask >> e & 1ull) ++k;
            return k;
        };
        std::vector<int> gstage(tasks.size(), -1);
        std::function<int(int)> stage_of_task = [&](int ti) -> int {
            if (gstage[ti] >= 0) return gstage[ti];
            int sg = 0;
            for (int tgt : deps[ti]) {
                if (tgt >= c->n_local) {
                    sg = std::max(sg, 1 + exch_index(std::min(63, std::max(0, c->cells[tgt].ps_level - m->ps_minlevel))));
                } else if (task_of[tgt] >= 0) {
                    sg = std::max(sg, stage_of_task(task_of[tgt]));
                }
            }
            return gstage[ti] = sg;
        };
        std::vector<char> in_gen(tasks.size(), 0);
        for (int ti : order)
            in_gen[ti] = !((by_level || wave[ti] == 0) && (is_regular(tasks[ti]) || is_regmapped(tasks[ti])));
        // With peers every launch list is split into tasks that can run before the df halo of the previous step has
        // arrived ("interior") and tasks that read a ghost cell's df, or project the slopes of a task that does
        // ("halo"): the interior part overlaps the transfer (the structure of flux!(p4est, ka), Flux.jl:461-485).
        std::vector<char> bnd(tasks.size(), 0);
        if (m->n_peer > 0) {
            const int g0 = c->n_local, g1 = c->n_local + c->n_ghost;
            for (int ti : order) {   // ascending wave: the tasks a task depends on come first
                bool b = false;
                for (int d = 0; d < D && !b; ++d)
                    for (int a = 0; a < tasks[ti].d[d].nA + tasks[ti].d[d].nB && !b; ++a) {
                        const int nc = cell_of_doff[c->slope_nb[tasks[ti].d[d].nb_begin + a].doff];
                        b = nc >= g0 && nc < g1;
                    }
                for (int tgt : deps[ti]) {
                    const int tj = (tgt < c->n_local) ? task_of[tgt] : -1;
                    if (tj >= 0 && bnd[tj]) b = true;
                }
                bnd[ti] = b ? 1 : 0;
            }
        }
        for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1)
            for (auto& kv : stages) {
                kv.second.reg_int = (int)kv.second.reg.size();
                kv.second.regm_int = (int)kv.second.regm.size();
                kv.second.gen_int = (int)kv.second.gen.size();
            }
        for (int ti : order) {
            if ((int)bnd[ti] != pass) continue;
            const int sidx = by_level ? stage_of_task(ti) : 0;
            kamr_ctx::SlopeStage& st = stages[sidx];
            st.wave = sidx;
            if (!in_gen[ti]) {
                if (is_regular(tasks[ti])) { st.reg.push_back(make_reg(tasks[ti])); continue; }
                SlopeRegMap rm;
                memset(&rm, 0, sizeof(rm));
                rm.r = make_reg(tasks[ti]);
                rm.goff = c->cells[tasks[ti].cell].goff;
                for (int d = 0; d < D; ++d)
                    for (int a = 0; a < 2; ++a) {
                        const SlopeNbr& e = c->slope_nb[tasks[ti].d[d].nb_begin + a];
                        rm.nb_rel[2 * d + a] = e.rel_off; rm.nb_goff[2 * d + a] = e.goff; rm.nb_np[2 * d + a] = e.np;
                    }
                st.regm.push_back(rm);
                continue;
            }
            SlopeTask t = tasks[ti];
            {   // dependencies computed by the same or an earlier launch of this sweep -> epoch flags
                t.dep_begin = (int)c->slope_deps.size();
                for (int tgt : deps[ti]) {
                    const int tj = (tgt < c->n_local) ? task_of[tgt] : -1;
                    if (tj >= 0 && in_gen[tj]) c->slope_deps.push_back(tgt);
                }
                t.dep_count = (int)c->slope_deps.size() - t.dep_begin;
                if (t.dep_count > NT_SLOPE) throw Fail("too many slope dependencies");
                st.flags = true;
            }
            st.gen.push_back(t);
        }
        }
        for (auto& kv : stages) {
            if (m->n_peer == 0) {
                kv.second.reg_int = (int)kv.second.reg.size();
                kv.second.regm_int = (int)kv.second.regm.size();
                kv.second.gen_int = (int)kv.second.gen.size();
            }
        }
        for (auto& kv : stages) {
            {   // prefetch distance of the regular launches: pf_mb of df ahead
                const double pf_mb = getenv("KAMR_PF_MB") ? atof(getenv("KAMR_PF_MB")) : 24.0;
                auto dist = [&](double pts, size_t cnt) {
                    if (!cnt || pf_mb <= 0.0) return 0;
                    return (int)std::min(512.0, std::max(1.0, pf_mb * 1048576.0 / (pts / cnt * 8.0 * K * (1 + D))));
                };
                double pr = 0, pm = 0;
                for (auto& r : kv.second.reg) pr += r.np;
                for (auto& r : kv.second.regm) pm += r.r.np;
                kv.second.pf_reg = dist(pr, kv.second.reg.size());
                kv.second.pf_regm = dist(pm, kv.second.regm.size());
            }
            kv.second.d_reg = c->dupload(kv.second.reg);
            kv.second.d_regm = c->dupload(kv.second.regm);
            kv.second.d_gen = c->dupload(kv.second.gen);
            c->slope_stages.push_back(std::move(kv.second));
        }
        c->dv.slope_deps = c->dupload(c->slope_deps);
        c->dv.slope_done = c->dalloc<int>((size_t)c->n_cell);
        CK(cudaMemsetAsync(c->dv.slope_done, 0, (size_t)c->n_cell * sizeof(int), c->stream));
        c->slope_epoch = 0;
        c->d_slope_ticket = c->dalloc<unsigned>(1);
        c->dv.err_flag = c->dalloc<int>(1);
        CK(cudaMemsetAsync(c->d_slope_ticket, 0, sizeof(unsigned), c->stream));
        CK(cudaMemsetAsync(c->dv.err_flag, 0, sizeof(int), c->stream));
        c->slope_ticket_base = 0;
    }
    lap("slope tasks");
    // ---- fluid cell list (Morton order: neighbours in space are neighbours in the launch, so the blocks
    // resident at one time share their neighbour reads through L2) and the phase-kernel bins
    for (int i = 0; i < c->n_local; ++i)
        if (c->cells[i].bound_enc >= 0) c->fluid_cells.push_back(i);
    c->d_fluid_cells = c->dupload(c->fluid_cells);
    c->limit_cells = c->fluid_cells;
    for (int i = c->n_local; i < c->n_local + c->n_ghost; ++i)
        if (c->cells[i].bound_enc >= 0) c->limit_cells.push_back(i);
    c->d_limit_cells = c->dupload(c->limit_cells);
    {
        // Launch classes.  A cell of n points is owned by C CTAs (a thread-block cluster), each with a contiguous range of
        // at most CHUNK_N points whose convected f (+ the h-plane of M[prim_c]) it stages in shared memory; giant cells
        // whose range does not fit even with C = 8 stage in the output array.  Ranges are grouped in steps of 256 points
        // so that a launch sizes its shared memory for what its cells need.
        const size_t smem_avail = (size_t)c->max_smem_optin - (14 << 10);   // static shared memory of the kernels
        std::map<std::vector<int>, Bin> bmap;
        c->fused_cells = 0;
        const double pf_mb = getenv("KAMR_PF_MB") ? atof(getenv("KAMR_PF_MB")) : 24.0;
        for (int cell : c->fluid_cells) {
            const CellInfo& ci = c->cells[cell];
            int C = 1;
            // staged planes per CTA: ~64 KB for the regular kernels (4 KB of static shared memory: 3 CTAs per SM), ~48 KB
            // for the general kernel (12 KB of slot / face records besides: at 64 KB only 2 CTAs fit, measured 9.65 vs
            // 8.5 ms on S4)
            const bool regular_cell = (ci.flags & (CELL_REGULAR | CELL_REGULAR_MAPPED)) != 0;
            const int chunk_n = (regular_cell ? CHUNK_BYTES : CHUNK_BYTES * 3 / 4) / (2 * K * (int)sizeof(double));
            while (C < 8 && (ci.n + C - 1) / C > (ci.n <= SMALL_N ? SMALL_N : chunk_n)) C *= 2;
            const int P = chunk_points(ci.n, C);
            const bool stage = sizeof(double) * ((size_t)(2 * K) * P + vtab_doubles(D, c->dv.n_vtab) + 2 +
                                                 (size_t)(D + 2) * batch_count(P)) <= smem_avail;
            const bool mapped = (ci.flags & CELL_REGULAR_MAPPED) != 0;
            const bool regular = mapped || (ci.flags & CELL_REGULAR) != 0;
            const int pclass = stage ? (P + 255) / 256 : 0;
            // with peers: cells that read ghost data (a ghost across a face) or wall data that depends on it (donors)
            // are launched after the halo has arrived, everything else while it is in flight
            bool halo = false;
            if (m->n_peer > 0) {
                halo = ci.bound_enc > 0;
                for (int q = ci.slot_begin; q < ci.slot_end && !halo; ++q) {
                    const int nb = c->slots[q].nbr;
                    halo = nb >= c->n_local && nb < c->n_local + c->n_ghost;
                }
            }
            Bin& b = bmap[{halo ? 1 : 0, regular ? (mapped ? 2 : 1) : 0, ci.n > SMALL_N ? 1 : 0, C, pclass}];
            b.halo = halo;
            b.C = C; b.wide = ci.n > SMALL_N; b.stage = stage; b.regular = regular; b.mapped = mapped;
            b.P = std::max(b.P, P);
            b.cells.push_back(cell);
            if (stage) c->fused_cells++;
        }
        std::vector<Bin> bins;
        // launch order: regular classes first (the wall kernels run beside them), then the rest
        for (int pass = 0; pass < 2; ++pass)
            for (auto& kv : bmap)
                if ((kv.first[1] != 0) == (pass == 0)) bins.push_back(std::move(kv.second));
        for (auto& b : bins) {   // DRAM -> L2 prefetch distance: pf_mb of cell state ahead in the launch
            double bytes = 0.0;
            for (int cell : b.cells) bytes += (double)c->cells[cell].np * 8.0 * K * (1 + D);
            bytes /= (double)b.cells.size();
            b.pf_dist = (int)std::min(512.0, std::max(1.0, pf_mb * 1048576.0 / bytes));
            if (pf_mb <= 0.0) b.pf_dist = 0;
        }
        for (auto& b : bins) {
            if (b.cells.empty()) continue;
            b.d_cells = c->dupload(b.cells);
            if (b.regular) {
                std::vector<RegCell> recs(b.cells.size());
                for (size_t q = 0; q < b.cells.size(); ++q) {
                    const CellInfo& ci = c->cells[b.cells[q]];
                    RegCell& r = recs[q];
                    memset(&r, 0, sizeof(r));
                    r.doff = ci.doff; r.goff = ci.goff; r.n = ci.n; r.np = ci.np; r.cell = b.cells[q]; r.vol = ci.vol;
                    for (int t = 0; t < D; ++t) r.mid[t] = ci.mid[t];
                    for (int k2 = 0; k2 < 2 * D; ++k2) {
                        const FaceRec& h = c->hot[ci.hot_begin + k2];
                        r.side[k2].ndoff = h.nf_off / K;
                        r.side[k2].area = h.area;
                        r.side[k2].fmid = h.fmid[k2 / 2];
                        r.side[k2].nmid = h.nbr_mid[k2 / 2];
                        r.side[k2].rel_off = (h.flags & 2) ? -1 : h.rel_off;
                        r.side[k2].ngoff = h.ngoff;
                        r.side[k2].np = h.np;
                    }
                }
                b.d_recs = c->dupload(recs);
                CK(cudaStreamSynchronize(c->stream));  // recs is a local
            }
            c->bins.push_back(b);
        }
    }
    lap("bins");
    // ---- immersed boundary tables (kernel d)
    std::vector<int> cvc_index;
    std::vector<double> cvc_gw, cvc_sw;
    if (m->ib) {
        const kamr_ib* ib = m->ib;
        if (ib->n_sn != m->n_solidnbr) throw Fail("kamr_ib.n_sn must equal kamr_mesh.n_solidnbr");
        auto add_nb = [&](int tgt, int src) {
            if (src < 0 || src >= c->n_cell) throw Fail("immersed-boundary neighbour id out of range");
            const CellInfo& cs = c->cells[src];
            IbNbr e;
            memset(&e, 0, sizeof(e));
            e.doff = cs.doff; e.goff = cs.goff; e.np = cs.np;
            e.rel_off = rel_offset(c, c->cells[tgt].grid, cs.grid);
            for (int t = 0; t < D; ++t) e.mid[t] = cs.mid[t];
            c->ib_nb.push_back(e);
        };
        for (int q = 0; q < ib->n_solid; ++q) {
            const int cell = ib->solid_cell[q];
            if (cell < 0 || cell >= c->n_local || c->cells[cell].bound_enc >= 0)
                throw Fail("solid_cell must be a local cell with bound_enc < 0");
            SolidTask t;
            memset(&t, 0, sizeof(t));
            t.cell = cell; t.nb_begin = (int)c->ib_nb.size();
            t.nb_count = ib->solid_nb_off[q + 1] - ib->solid_nb_off[q];
            if (t.nb_count <= 0 || t.nb_count > 32) throw Fail("solid cell needs 1..32 fluid neighbours");
            for (int a = ib->solid_nb_off[q]; a < ib->solid_nb_off[q + 1]; ++a) add_nb(cell, ib->solid_nb_ids[a]);
            c->solid_tasks.push_back(t);
        }
        const int sn0 = c->n_local + c->n_ghost;
        for (int q = 0; q < ib->n_sn; ++q) {
            SnTask t;
            memset(&t, 0, sizeof(t));
            t.sn_cell = sn0 + q; t.donor = ib->sn_donor[q]; t.solid = ib->sn_solid[q];
            if (t.donor < 0 || t.donor >= c->n_local) throw Fail("sn_donor must be a local cell");
            if (t.solid < 0 || t.solid >= sn0) throw Fail("sn_solid out of range");
            if (c->grid_canon[c->cells[t.sn_cell].grid] != c->grid_canon[c->cells[t.donor].grid]) throw Fail("a SolidNeighbor shares its donor's velocity grid");
            t.dir = ib->sn_faceid[q] / 2;
            t.nb_begin = (int)c->ib_nb.size();
            t.nb_count = ib->sn_nb_off[q + 1] - ib->sn_nb_off[q] + 1;
            if (t.nb_count > 8) throw Fail("too many fluid neighbours of a donor cell");
            for (int a = ib->sn_nb_off[q]; a < ib->sn_nb_off[q + 1]; ++a) add_nb(t.donor, ib->sn_nb_ids[a]);
            add_nb(t.donor, t.donor);  // fluid_cells[end] = ps_data, Immersed_boundary.jl:372
            t.cvc_begin = ib->cvc_off[q]; t.cvc_count = ib->cvc_off[q + 1] - ib->cvc_off[q];
            for (int a = t.cvc_begin + 1; a < t.cvc_begin + t.cvc_count; ++a)
                if (ib->cvc_index[a] <= ib->cvc_index[a - 1]) throw Fail("cvc_index must ascend within a SolidNeighbor");
            t.rel_ps = rel_offset(c, c->cells[t.donor].grid, c->cells[t.solid].grid);
            for (int k = 0; k < D; ++k) { t.aux[k] = ib->sn_aux[(size_t)q * D + k]; t.normal[k] = ib->sn_normal[(size_t)q * D + k]; }
            for (int k = 0; k < M; ++k) t.bc[k] = ib->sn_bc[(size_t)q * M + k];
            c->sn_tasks.push_back(t);
        }
        {   // per donor cell: its SolidNeighbor faces (positivity_preserving_ib!, CIP_Marching)
            std::vector<std::vector<int>> of_donor(c->n_local);
            for (int q = 0; q < ib->n_sn; ++q) of_donor[ib->sn_donor[q]].push_back(q);
            std::vector<DonorSn> dsn;
            for (int i = 0; i < c->n_local; ++i) {
                CellInfo& ci = c->cells[i];
                ci.sn_begin = (int)dsn.size();
                for (int q : of_donor[i]) {
                    DonorSn e;
                    memset(&e, 0, sizeof(e));
                    const CellInfo& sn = c->cells[sn0 + q];
                    e.doff = sn.doff; e.dir = ib->sn_faceid[q] / 2;
                    e.rot = (ib->sn_faceid[q] % 2 == 0) ? 1.0 : -1.0;   // get_rot, Theory/Math.jl:2
                    e.area = e.rot;
                    for (int t = 0; t < D; ++t) {
                        e.fmid[t] = ci.mid[t];
                        e.snmid[t] = sn.mid[t];
                        if (t != e.dir) e.area *= ci.ds[t];
                    }
                    e.fmid[e.dir] -= 0.5 * e.rot * ci.ds[e.dir];
                    dsn.push_back(e);
                }
                ci.sn_count = (int)dsn.size() - ci.sn_begin;
            }
            c->dv.donor_sn = c->dupload(dsn);
        }
        const int ncvc = ib->n_sn ? ib->cvc_off[ib->n_sn] : 0;
        cvc_index.assign(ib->cvc_index, ib->cvc_index + ncvc);
        cvc_gw.assign(ib->cvc_gas_w, ib->cvc_gas_w + ncvc);
        cvc_sw.assign(ib->cvc_solid_w, ib->cvc_solid_w + ncvc);
    } else if (m->n_solidnbr) {
        throw Fail("n_solidnbr > 0 needs the kamr_ib tables");
    }
    c->d_solid_tasks = c->dupload(c->solid_tasks);
    c->d_sn_tasks = c->dupload(c->sn_tasks);
    c->dv.ib_nb = c->dupload(c->ib_nb);
    c->dv.cvc_index = c->dupload(cvc_index);
    c->dv.cvc_gas_w = c->dupload(cvc_gw);
    c->dv.cvc_solid_w = c->dupload(cvc_sw);
    lap("ib tables");
    // ---- device state
    const size_t np = (size_t)c->npts_pad;
    c->dv.cells = c->dupload(c->cells);
    c->dv.slots = c->dupload(c->slots);
    c->dv.hot = c->dupload(c->hot);
    c->dv.rare = c->dupload(c->rare);
    c->dv.pm_start = c->dupload(c->pm_start);
    c->dv.slope_nb = c->dupload(c->slope_nb);
    c->dv.df = c->dalloc<double>(np * K);
    c->dv.df_new = c->dalloc<double>(np * K);
    c->dv.sdf = c->dalloc<double>(np * K * D);
    c->dv.sdl = c->dalloc<double>(np * K * D);
    c->dv.flux = nullptr;   // vs_data.flux exists only on the un-fused path: allocated on first use (ensure_flux)
    c->dv.w = c->dalloc<double>((size_t)c->n_cell * M);
    c->dv.prim = c->dalloc<double>((size_t)c->n_cell * M);
    c->dv.mflux = c->dalloc<double>((size_t)c->n_cell * M);
    c->dv.qf = c->dalloc<double>((size_t)c->n_cell * D);
    c->dv.sw = c->dalloc<double>((size_t)c->n_cell * M * D);
    c->dv.res_cell = c->dalloc<double>((size_t)c->n_local * 2 * M);
    c->dv.n_local = c->n_local;
    c->d_res = c->dalloc<double>(2 * M);
    CK(cudaMemsetAsync(c->dv.df, 0, np * K * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.df_new, 0, np * K * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.sdf, 0, np * K * D * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.sdl, 0, np * K * D * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.w, 0, (size_t)c->n_cell * M * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.prim, 0, (size_t)c->n_cell * M * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.mflux, 0, (size_t)c->n_cell * M * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.qf, 0, (size_t)c->n_cell * D * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.sw, 0, (size_t)c->n_cell * M * D * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->dv.res_cell, 0, (size_t)c->n_local * 2 * M * sizeof(double), c->stream));
    lap("device arrays");
    // ---- halo plan (Parallel/Ghost.jl:133-145, 203-284)
    c->halo_bytes_step = 0;
    long long send_total = 0, recv_total = 0;
    std::map<int, std::vector<int>> ghost_by_wave;
    for (int p = 0; p < m->n_peer; ++p) {
        PeerPlan pp;
        pp.rank = m->peer_rank[p];
        pp.send_base = send_total; pp.recv_base = recv_total;
        pp.mirror_first = m->send_off[p]; pp.mirror_count = m->send_off[p + 1] - m->send_off[p];
        pp.ghost_first = c->n_local + m->recv_off[p]; pp.ghost_count = m->recv_off[p + 1] - m->recv_off[p];
        long long pos = 0, spos_max = 0;
        std::map<int, long long> lpos;
        for (int q = m->send_off[p]; q < m->send_off[p + 1]; ++q) {
            const CellInfo& ci = c->cells[m->send_cells[q]];
            pp.df_send.push_back(CopySeg{ci.doff * K, pp.send_base + pos, (long long)ci.np * K});
            if (ci.bound_enc < 0) {  // solid ghost cells also travel mid-flux (solid_exchange_begin!)
                pp.solid.send.push_back(CopySeg{ci.doff * K, pp.send_base + pp.solid.send_len, (long long)ci.np * K});
                pp.solid.send_len += (long long)ci.np * K;
            }
            pos += (long long)ci.np * K;
            {   // sw of every mirror (fluid or not: the reference ships the block of every mirror, Ghost.jl:795-808)
                pp.sw.send.push_back(CopySeg{(long long)m->send_cells[q] * M * D, pp.send_base + pp.sw.send_len, (long long)M * D});
                pp.sw.send_len += (long long)M * D;
            }
            if (ci.bound_enc >= 0) {  // solid cells carry no slopes
                const int wv = std::max(0, ci.ps_level - m->ps_minlevel);
                auto& lv = pp.sdf[wv];
                long long& lp = lpos[wv];
                lv.send.push_back(CopySeg{ci.doff * K * D, pp.send_base + lp, (long long)ci.np * K * D});
                lp += (long long)ci.np * K * D;
                lv.send_len = lp;
                spos_max = std::max(spos_max, lp);
            }
        }
        pp.df_send_len = pos;
        const int g0 = c->n_local + m->recv_off[p], g1 = c->n_local + m->recv_off[p + 1];
        if (g1 > g0) {
            pp.df_recv_off = c->cells[g0].doff * K;
            pp.df_recv_len = (c->cells[g1 - 1].doff + c->cells[g1 - 1].np - c->cells[g0].doff) * K;
        }
        std::map<int, long long> rpos;
        long long rpos_max = 0;
        for (int gidx = g0; gidx < g1; ++gidx) {
            const CellInfo& ci = c->cells[gidx];
            pp.sw.recv.push_back(CopySeg{pp.recv_base + pp.sw.recv_len, (long long)gidx * M * D, (long long)M * D});
            pp.sw.recv_len += (long long)M * D;
            if (ci.bound_enc < 0) {
                pp.solid.recv.push_back(CopySeg{pp.recv_base + pp.solid.recv_len, ci.doff * K, (long long)ci.np * K});
                pp.solid.recv_len += (long long)ci.np * K;
                rpos_max = std::max(rpos_max, pp.solid.recv_len);
                continue;
            }
            const int wv = std::max(0, ci.ps_level - m->ps_minlevel);
            ghost_by_wave[wv].push_back(gidx);
            auto& lv = pp.sdf[wv];
            long long& rp = rpos[wv];
            lv.recv.push_back(CopySeg{pp.recv_base + rp, ci.doff * K * D, (long long)ci.np * K * D});
            rp += (long long)ci.np * K * D;
            lv.recv_len = rp;
            rpos_max = std::max(rpos_max, rp);
        }
        pp.early = (size_t)p < c->peer_early.size() ? c->peer_early[p] : ~0ull;
        for (auto& kv : pp.sdf) {
            kv.second.d_send = c->dupload(kv.second.send);
            kv.second.d_recv = c->dupload(kv.second.recv);
            c->halo_bytes_step += 8 * kv.second.send_len;
            if (pp.early >> std::min(63, kv.first) & 1ull) continue;
            // not exchanged early: appended (ascending level, mirror / ghost order inside a level) to the final message
            for (CopySeg sg : kv.second.send) {
                sg.dst = pp.send_base + pp.sdf_final.send_len;
                pp.sdf_final.send.push_back(sg);
                pp.sdf_final.send_len += sg.len;
            }
            for (CopySeg sg : kv.second.recv) {
                sg.src = pp.recv_base + pp.sdf_final.recv_len;
                pp.sdf_final.recv.push_back(sg);
                pp.sdf_final.recv_len += sg.len;
            }
        }
        pp.sdf_final.d_send = c->dupload(pp.sdf_final.send);
        pp.sdf_final.d_recv = c->dupload(pp.sdf_final.recv);
        spos_max = std::max(spos_max, std::max(pp.sdf_final.send_len, pp.sw.send_len));
        rpos_max = std::max(rpos_max, std::max(pp.sdf_final.recv_len, pp.sw.recv_len));
        pp.d_df_send = c->dupload(pp.df_send);
        pp.solid.d_send = c->dupload(pp.solid.send);
        pp.solid.d_recv = c->dupload(pp.solid.recv);
        c->halo_bytes_step += 8 * (pp.df_send_len + pp.solid.send_len);
        send_total += std::max(pos, spos_max);
        recv_total += rpos_max;
        c->peers.push_back(std::move(pp));
    }
    {
        std::vector<int> gf;
        for (auto& kv : ghost_by_wave) gf.insert(gf.end(), kv.second.begin(), kv.second.end());
        c->d_ghost_fluid = c->dupload(gf);
        c->n_ghost_fluid = (int)gf.size();
    }
    if (m->n_peer > 0) {
        c->d_sendbuf = c->dalloc<double>((size_t)send_total);
        c->d_recvbuf = c->dalloc<double>((size_t)recv_total);
        c->d_send_cells = c->dupload(std::vector<int>(m->send_cells, m->send_cells + m->send_off[m->n_peer]));
    }
    {   // neighbour lists as the host mesh holds them (PsData.neighbor), for the adaptation sensor
        const int ne = c->n_local * 2 * D;
        c->d_nb_state = c->dupload(std::vector<int>(m->nb_state, m->nb_state + ne));
        c->d_nb_off = c->dupload(std::vector<int>(m->nb_off, m->nb_off + ne + 1));
        c->d_nb_ids = c->dupload(std::vector<int>(m->nb_ids, m->nb_ids + m->nb_off[ne]));
        CK(cudaStreamSynchronize(c->stream));   // the temporaries above
    }
    lap("halo plan");
    setup_p2p(c, m);
    lap("one-sided halo set-up");
    if (getenv("KAMR_VERBOSE")) {
        for (auto& b : c->bins) {
            long long pts = 0;
            for (int cell : b.cells) pts += c->cells[cell].n;
            fprintf(stderr, "[kamr] phase bin: %zu cells, %lld points, C %d, P %d, %s, %s%s%s, prefetch %d cells\n",
                    b.cells.size(), pts, b.C, b.P, b.stage ? "smem" : "global", b.wide ? "256 thr " : "128 thr ",
                    b.regular ? "regular " : "general ", b.mapped ? "mapped" : "", b.pf_dist);
        }
        for (auto& st : c->slope_stages)
            fprintf(stderr, "[kamr] slope stage %d: %zu regular, %zu regular-mapped, %zu general%s\n", st.wave,
                    st.reg.size(), st.regm.size(), st.gen.size(), st.flags ? " (flags)" : "");
        fprintf(stderr, "[kamr] %zu solid cells, %zu solid neighbours, %zu pair maps (%zu ints)\n", c->solid_tasks.size(),
                c->sn_tasks.size(), c->rel_off.size(), c->pm_start.size());
    }
    CK(cudaStreamSynchronize(c->stream));
}

// ------------------------------------------------------------------------------------------------
// One-sided halo: at re-flatten time every pair of neighbouring ranks swaps the CUDA IPC handles of the arrays the
// other side writes into (both df allocations, the raw slopes, the flag table) and the block offsets of its ghost
// cells; a mirror cell's block is then STORED straight into the peer's ghost block over NVLink by a put kernel on a
// side stream, followed by a flag the receiver's stream waits on right before the first kernel that reads the ghosts.
// No rendezvous: a sender never waits, a receiver only when it reads.  p4est's mirror order of a pair equals the
// peer's ghost order (Parallel/Ghost.jl:221-264), so the k-th mirror sent to a peer lands in its k-th ghost from us.
#ifndef KAMR_PUT_CTAS
#define KAMR_PUT_CTAS 96
#endif
constexpr int PUT_CTAS = KAMR_PUT_CTAS;

void setup_p2p(kamr_ctx* c, const kamr_mesh* m) {
    if (m->n_peer == 0 || !c->comm) return;
    if (const char* h = getenv("KAMR_HALO")) if (!strcmp(h, "nccl")) return;   // two-sided NCCL path (debugging)
    const int K = c->K, D = c->D, np = m->n_peer, nr = c->cfg.nranks;
    P2P& pp = c->p2p;
    pp.d_flags = c->dalloc<int>((size_t)nr * HK_SLOTS);
    CK(cudaMemsetAsync(pp.d_flags, 0, sizeof(int) * (size_t)nr * HK_SLOTS, c->stream));
    constexpr int NH = 5;
    constexpr int HW = NH * (int)sizeof(cudaIpcMemHandle_t) / 8 + 1;   // handle block in 8-byte words + first ghost id
    cudaIpcMemHandle_t mine[NH];
    CK(cudaIpcGetMemHandle(&mine[0], c->dv.df));
    CK(cudaIpcGetMemHandle(&mine[1], c->dv.df_new));
    CK(cudaIpcGetMemHandle(&mine[2], c->dv.sdf));
    CK(cudaIpcGetMemHandle(&mine[3], pp.d_flags));
    CK(cudaIpcGetMemHandle(&mine[4], c->dv.sw));
    std::vector<long long> soff(np + 1, 0), roff(np + 1, 0);
    for (int p = 0; p < np; ++p) {
        soff[p + 1] = soff[p] + HW + (m->recv_off[p + 1] - m->recv_off[p]);   // what we tell p: handles + OUR ghost offsets
        roff[p + 1] = roff[p] + HW + (m->send_off[p + 1] - m->send_off[p]);   // what p tells us: handles + ITS ghost offsets
    }
    std::vector<long long> sx(soff[np]), rx(roff[np]);
    for (int p = 0; p < np; ++p) {
        memcpy(&sx[soff[p]], mine, sizeof(mine));
        sx[soff[p] + HW - 1] = c->n_local + m->recv_off[p];   // cell id of our first ghost from p (per-cell arrays)
        for (int gq = m->recv_off[p]; gq < m->recv_off[p + 1]; ++gq) sx[soff[p] + HW + gq - m->recv_off[p]] = c->cells[c->n_local + gq].doff;
    }
    long long* d_sx = c->dalloc<long long>(sx.size());
    long long* d_rx = c->dalloc<long long>(rx.size());
    CK(cudaMemcpyAsync(d_sx, sx.data(), sizeof(long long) * sx.size(), cudaMemcpyHostToDevice, c->stream));
    NCK(nccl().GroupStart());
    for (int p = 0; p < np; ++p) {
        NCK(nccl().Send(d_sx + soff[p], (size_t)(soff[p + 1] - soff[p]), ncclFloat64, m->peer_rank[p], c->comm, c->stream));
        NCK(nccl().Recv(d_rx + roff[p], (size_t)(roff[p + 1] - roff[p]), ncclFloat64, m->peer_rank[p], c->comm, c->stream));
    }
    NCK(nccl().GroupEnd());
    CK(cudaMemcpyAsync(rx.data(), d_rx, sizeof(long long) * rx.size(), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<double*> pdf0(np), pdf1(np), psdf(np), psw(np);
    std::vector<int*> pfl(np);
    for (int p = 0; p < np; ++p) {
        cudaIpcMemHandle_t h[NH];
        memcpy(h, &rx[roff[p]], sizeof(h));
        void* q[NH];
        for (int a = 0; a < NH; ++a) {
            cudaError_t e = cudaIpcOpenMemHandle(&q[a], h[a], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                throw Fail(std::string("cudaIpcOpenMemHandle (peer rank ") + std::to_string(m->peer_rank[p]) + "): " +
                           cudaGetErrorString(e) + " — set KAMR_HALO=nccl to use the two-sided NCCL halo instead");
            pp.opened.push_back(q[a]);
        }
        pdf0[p] = (double*)q[0]; pdf1[p] = (double*)q[1]; psdf[p] = (double*)q[2]; pfl[p] = (int*)q[3];
        psw[p] = (double*)q[4];
    }
    pp.d_pdf[0] = c->dupload(pdf0); pp.d_pdf[1] = c->dupload(pdf1);
    pp.d_psdf = c->dupload(psdf); pp.d_pflags = c->dupload(pfl); pp.d_psw = c->dupload(psw);
    // put lists and flag slots per message kind
    for (int p = 0; p < np; ++p) {
        const PeerPlan& pl = c->peers[p];
        const long long* rd = &rx[roff[p] + HW];
        auto kind_of_level = [&](int wv) { return (pl.early >> std::min(63, wv) & 1ull) ? HK_SDF_EARLY + std::min(63, wv) : HK_SDF_FINAL; };
        std::map<int, bool> sends, recvs;
        const long long rcell0 = rx[roff[p] + HW - 1];   // the peer's cell id of its first ghost from us
        const int MD = (D + 2) * D;
        for (int q = m->send_off[p]; q < m->send_off[p + 1]; ++q) {
            const CellInfo& ci = c->cells[m->send_cells[q]];
            const long long r = rd[q - m->send_off[p]];
            pp.msg[HK_SW].segs.push_back(PutSeg{(long long)m->send_cells[q] * MD, (rcell0 + q - m->send_off[p]) * MD, MD, p});
            sends[HK_SW] = true;
            pp.msg[HK_DF].segs.push_back(PutSeg{ci.doff * K, r * K, ci.np * K, p});
            sends[HK_DF] = true;
            if (ci.bound_enc < 0) {
                pp.msg[HK_SOLID].segs.push_back(PutSeg{ci.doff * K, r * K, ci.np * K, p});
                sends[HK_SOLID] = true;
            } else {
                const int kd = kind_of_level(std::max(0, ci.ps_level - m->ps_minlevel));
                pp.msg[kd].segs.push_back(PutSeg{ci.doff * K * D, r * K * D, ci.np * K * D, p});
                sends[kd] = true;
            }
        }
        for (int gq = m->recv_off[p]; gq < m->recv_off[p + 1]; ++gq) {
            const CellInfo& ci = c->cells[c->n_local + gq];
            recvs[HK_DF] = true;
            recvs[HK_SW] = true;
            if (ci.bound_enc < 0) recvs[HK_SOLID] = true;
            else recvs[kind_of_level(std::max(0, ci.ps_level - m->ps_minlevel))] = true;
        }
        for (auto& kv : sends) pp.msg[kv.first].to.push_back(p);
        for (auto& kv : recvs) pp.msg[kv.first].from.push_back(m->peer_rank[p] * HK_SLOTS + kv.first);
    }
    for (auto& kv : pp.msg) {
        kv.second.d_segs = c->dupload(kv.second.segs);
        kv.second.d_to = c->dupload(kv.second.to);
        kv.second.d_from = c->dupload(kv.second.from);
    }
    CK(cudaStreamSynchronize(c->stream));
    pp.on = true;
}

void exchange(kamr_ctx* c, int what, int level);

// the (what, level) of the two-sided path for a message kind
inline void kind_to_what(int kind, int& what, int& level) {
    level = 0;
    if (kind == HK_DF) what = 0;
    else if (kind == HK_SOLID) what = 2;
    else if (kind == HK_SDF_FINAL) what = 3;
    else if (kind == HK_SW) what = 4;
    else { what = 1; level = kind - HK_SDF_EARLY; }
}

// Send this rank's part of a message: the mirrors' blocks are stored into the peers' ghost blocks by a put kernel on
// the communication stream, then the flags go up.  Two-sided path: nothing yet (the exchange runs at the wait).
void halo_put(kamr_ctx* c, int kind, cudaStream_t after = nullptr) {
    if (c->peers.empty() || !c->p2p.on) return;
    if (!after) after = c->stream;
    auto it = c->p2p.msg.find(kind);
    if (it == c->p2p.msg.end()) return;
    HaloMsg& ms = it->second;
    ms.epoch++;
    if (ms.to.empty()) return;
    const bool is_df = kind == HK_DF || kind == HK_SOLID;
    CK(cudaEventRecord(c->ev_put_ready, after));
    CK(cudaStreamWaitEvent(c->comm_stream, c->ev_put_ready, 0));
    if (kind == HK_SW) {   // (D+2)*D doubles per cell: scalar stores
        Launch L_(c, KID_PACK, c->comm_stream);
        put_small_kernel<<<((int)ms.segs.size() + 7) / 8, 256, 0, c->comm_stream>>>(ms.d_segs, (int)ms.segs.size(), c->dv.sw,
                                                                                   c->p2p.d_psw);
    } else {
        Launch L_(c, KID_PACK, c->comm_stream);
        // (a modest grid: the NVLink stores need few CTAs to saturate the link, and every CTA of this high-priority
        // stream displaces one of the step's own kernels)
        put_segments_kernel<<<std::min<int>((int)ms.segs.size(), PUT_CTAS), 256, 0, c->comm_stream>>>(
            ms.d_segs, (int)ms.segs.size(), is_df ? c->dv.df : c->dv.sdf,
            is_df ? c->p2p.d_pdf[c->df_parity] : c->p2p.d_psdf);
    }
    halo_signal_kernel<<<1, (int)round_up((long long)ms.to.size(), 32), 0, c->comm_stream>>>(
        c->p2p.d_pflags, ms.d_to, (int)ms.to.size(), c->cfg.rank * HK_SLOTS + kind, ms.epoch);
    CK(cudaEventRecord(c->ev_put_done, c->comm_stream));
    c->put_in_flight = true;
    CK(cudaGetLastError());
}
// The main stream may not overwrite what a put in flight still reads (nor exit a step with one pending)
void halo_join_puts(kamr_ctx* c) {
    if (!c->put_in_flight) return;
    CK(cudaStreamWaitEvent(c->stream, c->ev_put_done, 0));
    c->put_in_flight = false;
}
// Before the first kernel that reads the ghosts of a message: wait for the peers' flags (one-sided) or run the
// pack / ncclSend+ncclRecv / unpack exchange here (two-sided).
void halo_wait(kamr_ctx* c, int kind, cudaStream_t on = nullptr) {
    if (c->peers.empty()) return;
    if (!on) on = c->stream;
    if (!c->p2p.on) {
        int what, level;
        kind_to_what(kind, what, level);
        exchange(c, what, level);
        return;
    }
    auto it = c->p2p.msg.find(kind);
    if (it == c->p2p.msg.end() || it->second.from.empty()) return;
    HaloMsg& ms = it->second;
    Launch L_(c, KID_UNPACK, on);
    halo_wait_kernel<<<1, (int)round_up((long long)ms.from.size(), 32), 0, on>>>(
        c->p2p.d_flags, ms.d_from, (int)ms.from.size(), ms.epoch, c->dv.err_flag);
    CK(cudaGetLastError());
}
// the df halo of the previous step is waited for lazily, right before the first kernel that reads a ghost's df
void halo_finish_df(kamr_ctx* c) {
    if (!c->df_wait_pending) return;
    c->df_wait_pending = false;
    halo_wait(c, HK_DF);
}

// vs_data.flux (and the SolidNeighbor "reconstruction perturbance" kept in the same array) is only needed by
// kamr_flux / kamr_iterate, CIP_Marching, Euler and the aux transfers; the fused CAIDVM step never touches it.
void ensure_flux(kamr_ctx* c) {
    if (c->dv.flux) return;
    const size_t n = (size_t)c->npts_pad * c->K;
    c->dv.flux = c->dalloc<double>(n);
    CK(cudaMemsetAsync(c->dv.flux, 0, n * sizeof(double), c->stream));
}

// Synchronises the stream and turns a device-side give-up (DevView::err_flag) into a host error.
void sync_and_check(kamr_ctx* c) {
    if (c->dv.err_flag) CK(cudaMemcpyAsync(c->h_err, c->dv.err_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->dv.err_flag && *c->h_err) {
        const int code = *c->h_err;
        *c->h_err = 0;
        CK(cudaMemsetAsync(c->dv.err_flag, 0, sizeof(int), c->stream));
        throw Fail(code == 2 ? "halo wait timed out: a peer never raised its flag (the results of this step are invalid)"
                             : "slope dependency sweep timed out waiting for another CTA (the results of this step are invalid)");
    }
}

// ------------------------------------------------------------------------------------------------
// host <-> device transfer of per-point arrays (host layout: unpadded planes; device: padded planes)
void copy_points(kamr_ctx* c, double* dev, double* host_rw, const double* host_ro, int comps, bool to_device) {
    const size_t total_h = (size_t)c->npts_host * comps;
    if (!c->padded) {
        if (to_device) CK(cudaMemcpyAsync(dev, host_ro, total_h * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        else CK(cudaMemcpyAsync(host_rw, dev, total_h * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return;
    }
    // padded: one contiguous transfer at full PCIe rate into / out of a device staging buffer, re-laid-out by a kernel
    if (c->d_stage_doubles < total_h) {
        if (c->d_stage) { CK(cudaStreamSynchronize(c->stream)); cudaFree(c->d_stage); }
        CK(cudaMalloc((void**)&c->d_stage, total_h * sizeof(double)));
        c->d_stage_doubles = total_h;
    }
    if (!c->d_host_off) c->d_host_off = c->dupload(c->host_off);
    const int grid = std::min(c->n_cell, 148 * 16);
    if (to_device) {
        CK(cudaMemcpyAsync(c->d_stage, host_ro, total_h * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        repack_kernel<<<grid, 256, 0, c->stream>>>(c->dv.cells, c->d_host_off, c->n_cell, comps, dev, c->d_stage, 1);
    } else {
        repack_kernel<<<grid, 256, 0, c->stream>>>(c->dv.cells, c->d_host_off, c->n_cell, comps, dev, c->d_stage, 0);
        CK(cudaMemcpyAsync(host_rw, c->d_stage, total_h * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
}

// ------------------------------------------------------------------------------------------------
// halo
PeerPlan::Lvl* halo_level(PeerPlan& pp, int what, int level) {
    if (what == 2) return (pp.solid.send.empty() && pp.solid.recv.empty()) ? nullptr : &pp.solid;
    if (what == 3) return (pp.sdf_final.send.empty() && pp.sdf_final.recv.empty()) ? nullptr : &pp.sdf_final;
    if (what == 4) return (pp.sw.send.empty() && pp.sw.recv.empty()) ? nullptr : &pp.sw;
    if (!(pp.early >> std::min(63, level) & 1ull)) return nullptr;   // this pair sends that level in the final message
    auto it = pp.sdf.find(level);
    return it == pp.sdf.end() ? nullptr : &it->second;
}

// what: 0 df of all mirrors (data_exchange!), 1 sdf of one level for the pairs that need it early
// (slope_exchange_level!), 2 df of solid ghost cells (solid_exchange_begin!/finish!), 3 sdf of every level not sent
// early, one message per pair, 4 macro slopes sw of all mirrors (sw_exchange!)
void exchange(kamr_ctx* c, int what, int level) {
    if (c->peers.empty()) return;
    if (!c->comm) throw Fail("mesh has peers but kamr_comm_init was not called");
    double* src = (what == 1 || what == 3) ? c->dv.sdf : (what == 4 ? c->dv.sw : c->dv.df);
    double* dst = src;
    // one pack and one unpack launch per exchange: the segment lists of all peers taking part, concatenated once
    auto key = std::make_pair(what, what == 1 ? level : 0);
    auto it = c->merged_segs.find(key);
    if (it == c->merged_segs.end()) {
        std::vector<CopySeg> snd, rcv;
        for (auto& pp : c->peers) {
            if (what == 0) {
                snd.insert(snd.end(), pp.df_send.begin(), pp.df_send.end());
            } else if (PeerPlan::Lvl* lv = halo_level(pp, what, level)) {
                snd.insert(snd.end(), lv->send.begin(), lv->send.end());
                rcv.insert(rcv.end(), lv->recv.begin(), lv->recv.end());
            }
        }
        kamr_ctx::MergedSegs ms;
        ms.n_send = (int)snd.size(); ms.n_recv = (int)rcv.size();
        ms.d_send = c->dupload(snd); ms.d_recv = c->dupload(rcv);
        CK(cudaStreamSynchronize(c->stream));   // snd / rcv are locals
        it = c->merged_segs.emplace(key, ms).first;
    }
    const kamr_ctx::MergedSegs& ms = it->second;
    bool any = false;
    for (auto& pp : c->peers) any = any || what == 0 || halo_level(pp, what, level) != nullptr;
    if (!any) return;
    if (ms.n_send) {
        Launch L_(c, KID_PACK);
        copy_segments_kernel<<<std::min<int>(ms.n_send, 2048), 256, 0, c->stream>>>(ms.d_send, ms.n_send, src,
                                                                                    c->d_sendbuf);
    }
    NCK(nccl().GroupStart());
    for (auto& pp : c->peers) {
        if (what == 0) {
            if (pp.df_send_len) NCK(nccl().Send(c->d_sendbuf + pp.send_base, (size_t)pp.df_send_len, ncclFloat64, pp.rank, c->comm, c->stream));
            if (pp.df_recv_len) NCK(nccl().Recv(c->dv.df + pp.df_recv_off, (size_t)pp.df_recv_len, ncclFloat64, pp.rank, c->comm, c->stream));
        } else {
            PeerPlan::Lvl* lv = halo_level(pp, what, level);
            if (!lv) continue;
            if (lv->send_len) NCK(nccl().Send(c->d_sendbuf + pp.send_base, (size_t)lv->send_len, ncclFloat64, pp.rank, c->comm, c->stream));
            if (lv->recv_len) NCK(nccl().Recv(c->d_recvbuf + pp.recv_base, (size_t)lv->recv_len, ncclFloat64, pp.rank, c->comm, c->stream));
        }
    }
    NCK(nccl().GroupEnd());
    if (ms.n_recv) {
        Launch L_(c, KID_UNPACK);
        copy_segments_kernel<<<std::min<int>(ms.n_recv, 2048), 256, 0, c->stream>>>(ms.d_recv, ms.n_recv, c->d_recvbuf,
                                                                                    dst);
    }
    CK(cudaGetLastError());
}

// Two-sided exchange of the mirrors' rows of a per-cell array (`width` doubles per cell, width <= (DIM+2)*DIM so that
// the sw message's staging region fits): gathered per peer, received straight into the peer's contiguous ghost rows.
// Event-driven users only (the adaptation sensor): the per-step halo is the one-sided path above.
void exchange_rows(kamr_ctx* c, double* arr, int width) {
    if (c->peers.empty()) return;
    if (!c->comm) throw Fail("mesh has peers but kamr_comm_init was not called");
    for (auto& pp : c->peers) {
        if (!pp.mirror_count) continue;
        Launch L_(c, KID_PACK);
        const long long total = (long long)pp.mirror_count * width;
        gather_rows_kernel<<<(int)std::min<long long>((total + 255) / 256, 1024), 256, 0, c->stream>>>(
            c->d_send_cells + pp.mirror_first, pp.mirror_count, width, arr, c->d_sendbuf + pp.send_base);
    }
    CK(cudaGetLastError());
    NCK(nccl().GroupStart());
    for (auto& pp : c->peers) {
        if (pp.mirror_count)
            NCK(nccl().Send(c->d_sendbuf + pp.send_base, (size_t)pp.mirror_count * width, ncclFloat64, pp.rank, c->comm, c->stream));
        if (pp.ghost_count)
            NCK(nccl().Recv(arr + (size_t)pp.ghost_first * width, (size_t)pp.ghost_count * width, ncclFloat64, pp.rank, c->comm, c->stream));
    }
    NCK(nccl().GroupEnd());
}

// update_criterion!(ka), Physical_space/AMR.jl:256-341.  The slopes are current (kamr_slope, as
// ps_adaptive_mesh_refinement! runs slope! first, AMR.jl:1109-1111): sw of local and ghost cells is on the device;
// the ghosts' w travels here (in the reference it rides in the df message, Parallel/Ghost.jl:1).
template <int D, int K>
void do_ps_criterion(kamr_ctx* c, double threshold, double* lohner_out, double* sensor_out) {
    constexpr int M = D + 2;
    if (!c->sw_valid)
        throw Fail("kamr_ps_criterion needs the macro slopes of the current state: call kamr_slope first "
                   "(ps_adaptive_mesh_refinement! runs slope! before update_criterion!)");
    halo_finish_df(c);
    halo_join_puts(c);
    if (!c->d_ps_lohner) {
        c->d_ps_lohner = c->dalloc<double>((size_t)c->n_local * M * D);
        c->d_ps_above = c->dalloc<double>((size_t)c->n_cell);
        c->d_ps_sensor = c->dalloc<double>((size_t)c->n_local);
        CK(cudaMemsetAsync(c->d_ps_above, 0, (size_t)c->n_cell * sizeof(double), c->stream));
    }
    exchange_rows(c, c->dv.w, M);
    const int n_real = c->n_local + c->n_ghost;
    const int grid = (c->n_local + 127) / 128;
    if (grid > 0) {
        Launch L_(c, KID_PS_SENSOR);
        ps_lohner_kernel<D><<<grid, 128, 0, c->stream>>>(c->dv.cells, c->d_nb_state, c->d_nb_off, c->d_nb_ids, c->dv.w,
                                                         c->dv.prim, c->dv.sw, c->n_local, n_real, c->gas.gamma, threshold,
                                                         c->d_ps_lohner, c->d_ps_above);
    }
    exchange_rows(c, c->d_ps_above, 1);   // lohner_flag_exchange!, Parallel/Ghost.jl:939-978
    if (grid > 0) {
        Launch L_(c, KID_PS_SENSOR);
        ps_buffer_kernel<D><<<grid, 128, 0, c->stream>>>(c->dv.cells, c->d_nb_state, c->d_nb_off, c->d_nb_ids, c->n_local,
                                                         n_real, threshold, c->d_ps_above, c->d_ps_lohner, c->d_ps_sensor);
    }
    CK(cudaGetLastError());
    if (lohner_out)
        CK(cudaMemcpyAsync(lohner_out, c->d_ps_lohner, (size_t)c->n_local * M * D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (sensor_out)
        CK(cudaMemcpyAsync(sensor_out, c->d_ps_sensor, (size_t)c->n_local * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    sync_and_check(c);
}

// ---- velocity-space adaptation inputs
VsPar make_vs_par(kamr_ctx* c, const kamr_vs_adapt* p) {
    if (!p) throw Fail("kamr_vs_adapt is NULL");
    if (p->mode != 0 && p->mode != 1) throw Fail("kamr_vs_adapt.mode: 0 (:lohner) or 1 (contribution)");
    if (p->maxlevel < 0 || p->maxlevel >= VPK_LEVELS) throw Fail("kamr_vs_adapt.maxlevel out of range");
    VsPar v{};
    v.mode = p->mode; v.maxlevel = p->maxlevel;
    double du = 1.0, nt = 1.0;
    for (int d = 0; d < c->D; ++d) {
        if (p->trees[d] <= 0 || !(p->vmax[d] > p->vmin[d])) throw Fail("kamr_vs_adapt: trees / quadrature invalid");
        const double ds0 = (p->vmax[d] - p->vmin[d]) / p->trees[d];   // Velocity_space/AMR.jl:29-30
        v.vmin[d] = p->vmin[d];
        v.h_fine[d] = ds0 / std::ldexp(1.0, p->maxlevel);             // Neighbor.jl:83
        const long long gm = (long long)p->trees[d] << p->maxlevel;
        if (gm > (1 << 20)) throw Fail("kamr_vs_adapt: finest-level velocity lattice too large");
        v.gmax[d] = (int)gm;
        du *= p->vmax[d] - p->vmin[d];
        nt *= (double)p->trees[d];
    }
    v.coeff_lohner = p->coeff_lohner; v.coeff_local = p->coeff_local; v.coeff_global = p->coeff_global;
    v.vr_density = p->vr_density; v.vr_energy = p->vr_energy;
    v.cell_weight = du / nt / std::ldexp(1.0, c->D * p->maxlevel);    // AMR.jl:163
    return v;
}

// face-neighbour tables of every distinct velocity grid, built on the device (one CTA per grid, a finest-level
// lattice per CTA in a scratch buffer, grids taken in batches that keep the scratch under 256 MB)
template <int D, int K>
void ensure_vs_neighbors(kamr_ctx* c, const kamr_vs_adapt* p, const VsPar& vp) {
    auto& vc = c->vs_cache;
    bool same = vc.valid && vc.maxlevel == p->maxlevel;
    for (int d = 0; d < D && same; ++d)
        same = vc.trees[d] == p->trees[d] && vc.vmin[d] == p->vmin[d] && vc.vmax[d] == p->vmax[d];
    if (same) return;
    const int ng = (int)c->grid_n.size();
    std::vector<VsGridTask> tasks;
    std::vector<long long> nb_off(ng, 0);
    long long total = 0;
    for (int g = 0; g < ng; ++g) {
        if (c->grid_canon[g] != g) continue;
        nb_off[g] = total;
        tasks.push_back(VsGridTask{c->grid_goff[g], total, c->grid_n[g], c->grid_np[g]});
        total += (long long)c->grid_n[g] * D * 2;
    }
    for (int g = 0; g < ng; ++g) nb_off[g] = nb_off[c->grid_canon[g]];
    vc.d_nbt = c->dalloc<int>((size_t)total);
    vc.d_nb_off = c->dupload(nb_off);
    VsGridTask* d_tasks = c->dupload(tasks);
    long long lattice = 1;
    for (int d = 0; d < D; ++d) lattice *= vp.gmax[d];
    const int batch = (int)std::max<long long>(1, std::min<long long>((long long)tasks.size(), (256ll << 20) / (lattice * 4)));
    int* d_lattice = nullptr;
    CK(cudaMalloc((void**)&d_lattice, (size_t)batch * lattice * sizeof(int)));
    for (size_t first = 0; first < tasks.size(); first += batch) {
        const int nb = (int)std::min<size_t>(batch, tasks.size() - first);
        Launch L_(c, KID_PS_SENSOR);
        vs_neighbors_kernel<D><<<nb, 256, 0, c->stream>>>(d_tasks + first, vp, c->dv.v_mid, c->dv.v_level, d_lattice,
                                                          lattice, vc.d_nbt, c->dv.err_flag);
    }
    cudaError_t e = cudaGetLastError();
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_lattice);
    CK(e);
    CK(cudaMemcpy(c->h_err, c->dv.err_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (*c->h_err == 3) {
        *c->h_err = 0;
        CK(cudaMemset(c->dv.err_flag, 0, sizeof(int)));
        throw Fail("kamr_vs_adapt does not describe the velocity grids on the device (a velocity cell lies outside "
                   "quadrature / vs_trees_num / AMR_VS_MAXLEVEL)");
    }
    vc.valid = true; vc.maxlevel = p->maxlevel;
    for (int d = 0; d < D; ++d) { vc.trees[d] = p->trees[d]; vc.vmin[d] = p->vmin[d]; vc.vmax[d] = p->vmax[d]; }
}

template <int D, int K>
void do_vs_criterion(kamr_ctx* c, const kamr_vs_adapt* p, uint8_t* refine_flag, uint8_t* coarsen_ok) {
    const VsPar vp = make_vs_par(c, p);
    if (!c->raw_sdf_valid)
        throw Fail("kamr_vs_criterion reads the raw slopes (criterion distribution df + max_d |sdf ds_d|): none are "
                   "resident — call kamr_slope, or set KAMR_OPT_KEEP_SDF before the last step");
    if (c->n_local == 0) return;
    halo_finish_df(c);
    halo_join_puts(c);
    if (vp.mode == 0) ensure_vs_neighbors<D, K>(c, p, vp);
    const long long npts = c->host_off[c->n_local];
    if (!c->d_vs_flags) c->d_vs_flags = c->dalloc<unsigned char>((size_t)2 * npts);
    if (!c->d_host_off) c->d_host_off = c->dupload(c->host_off);
    {
        Launch L_(c, KID_PS_SENSOR);
        vs_criterion_kernel<D, K><<<c->n_local, 256, 0, c->stream>>>(c->dv, vp, c->d_host_off, c->vs_cache.d_nb_off,
                                                                     c->vs_cache.d_nbt, c->d_vs_flags, c->d_vs_flags + npts);
    }
    CK(cudaGetLastError());
    if (refine_flag) CK(cudaMemcpyAsync(refine_flag, c->d_vs_flags, (size_t)npts, cudaMemcpyDeviceToHost, c->stream));
    if (coarsen_ok) CK(cudaMemcpyAsync(coarsen_ok, c->d_vs_flags + npts, (size_t)npts, cudaMemcpyDeviceToHost, c->stream));
    sync_and_check(c);
}

template <int D, int K>
void do_vs_resolution(kamr_ctx* c, const kamr_vs_adapt* p, double* out) {
    const VsPar vp = make_vs_par(c, p);
    out[0] = out[1] = 0.0;
    if (c->n_local == 0) return;
    halo_finish_df(c);
    halo_join_puts(c);
    if (!c->d_vs_res) c->d_vs_res = c->dalloc<double>((size_t)2 * c->n_local);
    {
        Launch L_(c, KID_PS_SENSOR);
        vs_resolution_kernel<D, K><<<c->n_local, 256, 0, c->stream>>>(c->dv, vp, c->d_vs_res);
    }
    CK(cudaGetLastError());
    std::vector<double> h((size_t)2 * c->n_local);
    CK(cudaMemcpyAsync(h.data(), c->d_vs_res, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    sync_and_check(c);
    for (int i = 0; i < c->n_local; ++i) {   // density_res = max(density_res_i, density_res), AMR.jl:146-147
        out[0] = std::max(h[2 * i], out[0]);
        out[1] = std::max(h[2 * i + 1], out[1]);
    }
}

template <int D, int K>
void do_project_cells(kamr_ctx* c, int n, const int32_t* cells) {
    if (n <= 0) return;
    if (!cells) throw Fail("kamr_project_cells: cells is NULL");
    std::vector<int> list;
    for (int q = 0; q < n; ++q) {
        if (cells[q] < 0 || cells[q] >= c->n_local) throw Fail("kamr_project_cells: cell id outside [0, n_local)");
        if (c->cells[cells[q]].bound_enc >= 0) list.push_back(cells[q]);   // AMR.jl:128
    }
    if (list.empty()) return;
    halo_finish_df(c);
    halo_join_puts(c);
    int* d_list = nullptr;
    CK(cudaMalloc((void**)&d_list, list.size() * sizeof(int)));
    CK(cudaMemcpyAsync(d_list, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    {
        Launch L_(c, KID_UPDATE);
        project_cells_kernel<D, K><<<(int)list.size(), 256, 0, c->stream>>>(c->dv, d_list);
    }
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(c->stream);
    cudaFree(d_list);
    CK(e);
    c->sw_valid = false; c->raw_sdf_valid = false;
}

// ------------------------------------------------------------------------------------------------
// launch sequences
#ifndef KAMR_NT
#define KAMR_NT 256
#endif
#ifndef KAMR_PNT
#define KAMR_PNT 128
#endif
#ifndef KAMR_MINB
#define KAMR_MINB 6
#endif
constexpr int NT = KAMR_NT;      // threads per CTA of the slope kernel (one CTA per physical cell)
constexpr int PNT = KAMR_PNT;
// threads per CTA of the phase kernel for small cells: small CTAs keep many cells in flight per SM, so one cell's barriers
// and serial moments->prim step hide behind the others
#ifndef KAMR_MINB_GEN
#define KAMR_MINB_GEN KAMR_MINB
#endif
constexpr int MINB_GEN = KAMR_MINB_GEN;  // same for the general phase kernel (more live state: slot loops, pair-mapped gather)
constexpr int MINB = KAMR_MINB;  // CTAs per SM the phase kernel is register-budgeted for

// part 0: the tasks that read no ghost data, part 1: the rest, part 2: both
template <int D, int K>
void launch_slope_stage(kamr_ctx* c, const kamr_ctx::SlopeStage& st, int raw_all, int part = 2) {
    auto range = [&](int total, int n_int, int& first, int& count) {
        first = part == 1 ? n_int : 0;
        count = part == 0 ? n_int : (part == 1 ? total - n_int : total);
    };
    int f0, n0;
    range((int)st.reg.size(), st.reg_int, f0, n0);
    if (n0 > 0) {
        Launch L_(c, KID_SLOPE_REGULAR);
        slope_regular_kernel<D, K, NT, false><<<n0, NT, 0, c->stream>>>(c->dv, st.d_reg + f0, raw_all, st.pf_reg);
    }
    range((int)st.regm.size(), st.regm_int, f0, n0);
    if (n0 > 0) {
        Launch L_(c, KID_SLOPE_REGMAP);
        slope_regular_kernel<D, K, NT, true><<<n0, NT, 0, c->stream>>>(c->dv, st.d_regm + f0, raw_all, st.pf_regm);
    }
    range((int)st.gen.size(), st.gen_int, f0, n0);
    if (n0 > 0) {
        Launch L_(c, KID_SLOPE);
        slope_kernel<D, K, true, NT_SLOPE><<<n0, NT_SLOPE, 0, c->stream>>>(
            c->dv, st.d_gen + f0, raw_all, st.flags ? c->slope_epoch : 0, c->d_slope_ticket, c->slope_ticket_base);
        if (st.flags) c->slope_ticket_base += (unsigned)n0;   // tickets drawn by this launch (wraps with the counter)
    }
}

template <int D, int K>
void run_limit(kamr_ctx* c, const int* d_cells, int n) {
    if (!n) return;
    Launch L_(c, KID_LIMIT);
    limit_kernel<D, K><<<n, 256, 0, c->stream>>>(c->dv, d_cells);
}

template <int D, int K>
void run_macro_slope(kamr_ctx* c) {
    if (c->fluid_cells.empty()) return;
    Launch L_(c, KID_MACRO_SLOPE);
    macro_slope_kernel<D, K><<<(int)c->fluid_cells.size(), 256, 0, c->stream>>>(c->dv, c->d_fluid_cells);
}

// the slope halo of the mirrors not exchanged early has been sent; wait for the peers' and limit the ghosts' slopes
template <int D, int K>
void finish_slope_halo(kamr_ctx* c) {
    if (c->peers.empty()) return;
    halo_wait(c, HK_SDF_FINAL);
    run_limit<D, K>(c, c->d_ghost_fluid, c->n_ghost_fluid);
}

// slope!(p4est, ka).  raw_all: every cell's reference sdf is written (public kamr_slope, KAMR_OPT_KEEP_SDF);
// otherwise only where a kernel reads it, and the limited slopes everywhere.
// With peers: the tasks that read no ghost data are launched first; then the stream waits for the df halo of the
// previous step, the remaining tasks follow, and the mirrors' slopes are put.  defer_final: the caller overlaps the
// final slope message with work that reads no ghost slopes and calls finish_slope_halo itself.
template <int D, int K>
void do_slope(kamr_ctx* c, bool with_sw, bool raw_all, bool defer_final = false) {
    raw_all = raw_all || c->keep_sdf;
    c->slope_epoch = c->slope_epoch == 0x7fffffff ? 1 : c->slope_epoch + 1;
    if (c->peers.empty()) {
        for (auto& st : c->slope_stages) launch_slope_stage<D, K>(c, st, raw_all);
    } else {
        halo_join_puts(c);   // (the slope kernels overwrite what the previous step's slope puts read)
        // stage s = the waves up to and including the s-th early wave; it is followed by that wave's exchange between
        // the pairs that project each other's slopes of that level (slope_exchange_level!, Parallel/Ghost.jl:896).
        // Everything else travels in one message per pair after the last stage, then the ghosts' limited slopes.
        std::vector<int> early;
        for (int e = 0; e < 64; ++e)
            if (c->early_mask >> e & 1ull) early.push_back(e);
        size_t nst = early.size() + 1;
        for (auto& st : c->slope_stages) nst = std::max(nst, (size_t)st.wave + 1);
        for (size_t sidx = 0; sidx < nst; ++sidx) {
            for (auto& st : c->slope_stages)
                if ((size_t)st.wave == sidx) launch_slope_stage<D, K>(c, st, raw_all, 0);
            if (sidx == 0) halo_finish_df(c);
            for (auto& st : c->slope_stages)
                if ((size_t)st.wave == sidx) launch_slope_stage<D, K>(c, st, raw_all, 1);
            if (sidx < early.size()) {
                halo_put(c, HK_SDF_EARLY + early[sidx]);
                halo_wait(c, HK_SDF_EARLY + early[sidx]);
            }
        }
        halo_put(c, HK_SDF_FINAL);
        if (!defer_final) finish_slope_halo<D, K>(c);
    }
    c->raw_sdf_valid = raw_all;
    CK(cudaGetLastError());
    c->sw_valid = with_sw;
    if (with_sw) {
        run_macro_slope<D, K>(c);
        if (!c->peers.empty()) {   // sw_exchange!, Parallel/Ghost.jl:867: the ghosts' sw feed the host's Löhner sensor
            halo_join_puts(c);
            halo_put(c, HK_SW);
            halo_wait(c, HK_SW);
        }
    }
    CK(cudaGetLastError());
}

template <class Kern>
void prepare_kernel(Kern kern, int max_dyn) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    // no carve-out preference: the driver sizes shared memory for the resident CTAs and leaves the rest of the 256 KB to
    // L1 (forcing MaxShared cost 12 % of the step, r01_i)
}

// launches `kern` on ncell cells with C CTAs per cell (a thread-block cluster when C > 1)
template <class Kern, class... Args>
void launch_cells(kamr_ctx* c, Kern kern, int ncell, int C, int nt, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ncell * C, 1, 1);
    cfg.blockDim = dim3(nt, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = C > 1 ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, args...));
}

#ifndef KAMR_MINB_WIDE
#define KAMR_MINB_WIDE 3
#endif
constexpr int PNT_WIDE = 256;              // threads per CTA for point ranges of more than SMALL_N points
constexpr int MINB_WIDE = KAMR_MINB_WIDE;  // CTAs per SM those kernels are register-budgeted for

template <int D, int K, int MODE, bool STAGE, int PT, int MB, int C>
void launch_phase_inst(kamr_ctx* c, const Bin& b, size_t smem, double dt, int want, int kid) {
    auto kern = phase_kernel<D, K, MODE, STAGE, PT, MB, C>;
    static bool prepared = false;  // per instantiation; attributes are per device function
    if (!prepared) {
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, kern));
        prepare_kernel(kern, c->max_smem_optin - (int)fa.sharedSizeBytes);
        prepared = true;
    }
    Launch L_(c, kid);
    launch_cells(c, kern, (int)b.cells.size(), C, PT, smem, c->dv, c->gas, (const int*)b.d_cells, dt, want, b.pf_dist);
}

template <int D, int K, bool STAGE, int PT, int MB, bool MAPPED, int C>
void launch_regular_inst2(kamr_ctx* c, const Bin& b, size_t smem, double dt, int want) {
    auto kern = phase_regular_kernel<D, K, STAGE, PT, MB, MAPPED, C>;
    static bool prepared = false;
    if (!prepared) {
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, kern));
        prepare_kernel(kern, c->max_smem_optin - (int)fa.sharedSizeBytes);
        prepared = true;
    }
    Launch L_(c, MAPPED ? KID_STEP_REGMAP : KID_STEP_REGULAR);
    launch_cells(c, kern, (int)b.cells.size(), C, PT, smem, c->dv, c->gas, (const RegCell*)b.d_recs, dt, want, b.pf_dist);
}
#ifndef KAMR_PAIR_THREADS
#define KAMR_PAIR_THREADS 512
#endif
template <int D, int K, bool STAGE, int PT, int MB, int C>
void launch_regular_inst(kamr_ctx* c, const Bin& b, size_t smem, double dt, int want) {
    // -DKAMR_PAIRS: the plain regular kernel takes two points per thread (128-bit loads): twice the live values, so it
    // is register-budgeted for KAMR_PAIR_THREADS resident threads per SM (128 registers at 512) instead of 768
#ifdef KAMR_PAIRS
    constexpr int MBP = KAMR_PAIR_THREADS / PT;
#else
    constexpr int MBP = MB;
#endif
    if (b.mapped) launch_regular_inst2<D, K, STAGE, PT, MB, true, C>(c, b, smem, dt, want);
    else launch_regular_inst2<D, K, STAGE, PT, MBP, false, C>(c, b, smem, dt, want);
}

// the fused step of one bin: the instantiation follows the bin's CTA width, cluster size and staging area
template <int D, int K>
void launch_fused(kamr_ctx* c, const Bin& b, double dt, int want) {
    const size_t smem = phase_smem_bytes<D, K>(c->dv.n_vtab, b.P, b.stage, b.wide || !b.regular, !b.regular || b.mapped);
#define KAMR_FUSED(STAGE, PT, MB, CC)                                                                       \
    do {                                                                                                    \
        if (b.regular) launch_regular_inst<D, K, STAGE, PT, MB, CC>(c, b, smem, dt, want);                  \
        else launch_phase_inst<D, K, MODE_FUSED, STAGE, PT, MB, CC>(c, b, smem, dt, want, KID_STEP);        \
    } while (0)
    if (!b.wide) { KAMR_FUSED(true, PNT, MINB, 1); return; }
    if (!b.stage) { KAMR_FUSED(false, PNT_WIDE, MINB_WIDE, 8); return; }
    switch (b.C) {
        case 1: KAMR_FUSED(true, PNT_WIDE, MINB_WIDE, 1); break;
        case 2: KAMR_FUSED(true, PNT_WIDE, MINB_WIDE, 2); break;
        case 4: KAMR_FUSED(true, PNT_WIDE, MINB_WIDE, 4); break;
        default: KAMR_FUSED(true, PNT_WIDE, MINB_WIDE, 8); break;
    }
#undef KAMR_FUSED
}

// flux!(p4est, ka) / iterate!(CAIDVM_Marching) one call at a time: one CTA per cell over all its points; the update
// stages in shared memory when the whole cell fits
template <int D, int K, int MODE>
void launch_unfused(kamr_ctx* c, const Bin& b, double dt, int want) {
    const int kid = MODE == MODE_FLUX ? KID_FLUX : KID_UPDATE;
    int nmax = 0;
    for (int cell : b.cells) nmax = std::max(nmax, c->cells[cell].n);
    const int P = chunk_points(nmax, 1);
    const bool stage = MODE == MODE_UPDATE &&
                       phase_smem_bytes<D, K>(c->dv.n_vtab, P, true, false, false) + (14 << 10) <= (size_t)c->max_smem_optin;
    const size_t smem = phase_smem_bytes<D, K>(c->dv.n_vtab, P, stage, false, false);
    if (stage) launch_phase_inst<D, K, MODE, true, PNT_WIDE, MINB_WIDE, 1>(c, b, smem, dt, want, kid);
    else launch_phase_inst<D, K, MODE, false, PNT_WIDE, MINB_WIDE, 1>(c, b, smem, dt, want, kid);
}

// the wall half of flux!(p4est, ka) (Flux.jl:463-481): update_solid_cell!, the solid halo, update_solid_neighbor!.
// df2: second buffer that receives the same values (the write side of a fused step), or null.
template <int D, int K>
void do_ib(kamr_ctx* c, double* df2, cudaStream_t st = nullptr) {
    if (!st) st = c->stream;
    if (!c->solid_tasks.empty()) {
        Launch L_(c, KID_SOLID_CELL, st);
        solid_cell_kernel<D, K><<<(int)c->solid_tasks.size(), 256, 0, st>>>(c->dv, c->gas, c->d_solid_tasks, df2);
    }
    if (st == c->stream || c->p2p.on) {   // (the two-sided exchange lives on the main stream)
        halo_put(c, HK_SOLID, st);
        halo_wait(c, HK_SOLID, st);
    }
    if (!c->sn_tasks.empty()) {
        Launch L_(c, KID_SOLID_NBR, st);
        solid_neighbor_kernel<D, K><<<(int)c->sn_tasks.size(), 256, 0, st>>>(c->dv, c->gas, c->d_sn_tasks, df2);
    }
    CK(cudaGetLastError());
}

template <int D, int K>
void do_flux(kamr_ctx* c, double dt) {
    ensure_flux(c);
    halo_finish_df(c);
    do_ib<D, K>(c, nullptr);
    for (auto& b : c->bins) launch_unfused<D, K, MODE_FLUX>(c, b, dt, 0);
    CK(cudaGetLastError());
}

template <int D, int K>
void fetch_residual(kamr_ctx* c, int want, double* res_out) {
    if (!want) return;
    const int M = D + 2;
    { Launch L_(c, KID_RESIDUAL);
      residual_reduce_kernel<<<2 * M, 1024, 0, c->stream>>>(c->dv.res_cell, c->d_fluid_cells, (int)c->fluid_cells.size(),
                                                       2 * M, c->d_res); }
    CK(cudaMemcpyAsync(c->h_res, c->d_res, 2 * M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    sync_and_check(c);
    if (res_out) memcpy(res_out, c->h_res, 2 * M * sizeof(double));
}

// data_exchange! (Parallel/Ghost.jl:841) after the update: the mirrors' new df goes out now; its arrival is waited for
// by the first kernel of the next call that reads a ghost's df
void send_df_halo(kamr_ctx* c) {
    if (c->peers.empty()) return;
    halo_join_puts(c);
    halo_put(c, HK_DF);
    c->df_wait_pending = true;
}

template <int D, int K>
void do_iterate(kamr_ctx* c, double dt, int want, double* res_out) {
    c->sw_valid = false;
    ensure_flux(c);
    halo_finish_df(c);
    if (c->gas.marching == KAMR_MARCH_CIP) {
        if (!c->fluid_cells.empty()) {
            Launch L_(c, KID_UPDATE);
            cip_update_kernel<D, K><<<(int)c->fluid_cells.size(), 256, 0, c->stream>>>(c->dv, c->gas, c->d_fluid_cells,
                                                                                        dt, want);
        }
    } else if (c->gas.marching == KAMR_MARCH_EULER) {
        if (!c->fluid_cells.empty()) {
            Launch L_(c, KID_UPDATE);
            euler_update_kernel<D, K><<<(int)c->fluid_cells.size(), 256, 0, c->stream>>>(c->dv, c->gas, c->d_fluid_cells,
                                                                                          dt, want);
        }
    } else {
        for (auto& b : c->bins) launch_unfused<D, K, MODE_UPDATE>(c, b, dt, want);
    }
    CK(cudaGetLastError());
    // The update ran IN PLACE: unlike the fused step, whose ghosts of the new time level land in the other df buffer, a
    // one-sided put here could overwrite ghost blocks a slower peer is still reading.  data_exchange! therefore stays a
    // two-sided pack / ncclSend+ncclRecv on this path (a rendezvous at the same program point of both ranks).
    if (!c->peers.empty()) {
        halo_join_puts(c);
        exchange(c, 0, 0);
    }
    fetch_residual<D, K>(c, want, res_out);
}

template <int D, int K>
void do_step(kamr_ctx* c, double dt, int want, double* res_out) {
    c->sw_valid = false;
    if (c->gas.marching != KAMR_MARCH_CAIDVM) {  // only CAIDVM_Marching has the fused kernel
        do_slope<D, K>(c, false, false);
        do_flux<D, K>(c, dt);
        do_iterate<D, K>(c, dt, want, res_out);
        return;
    }
    if (c->peers.empty()) {
        do_slope<D, K>(c, false, false);
        // The wall kernels feed the donor cells only, and a donor is never a regular cell: they run on the side stream
        // while the regular cells' phase kernels run here (the overlap flux!(p4est, ka) has between the solid halo and
        // the non-IB faces, Flux.jl:461-485).
        const bool side = !c->solid_tasks.empty() || !c->sn_tasks.empty();
        if (side) {
            CK(cudaEventRecord(c->ev_fork, c->stream));
            CK(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
            do_ib<D, K>(c, c->dv.df_new, c->side_stream);
            CK(cudaEventRecord(c->ev_join, c->side_stream));
        }
        for (auto& b : c->bins) if (b.regular) launch_fused<D, K>(c, b, dt, want);
        if (side) CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        for (auto& b : c->bins) if (!b.regular) launch_fused<D, K>(c, b, dt, want);
    } else {
        // With peers (the order of flux!(p4est, ka), Flux.jl:461-485, with more of it overlapped): slopes; the final
        // slope message goes out; the cells that read no ghost data are updated while it travels; then the ghosts'
        // slopes, the wall kernels around the solid-cell halo, and the cells along the partition boundary.
        do_slope<D, K>(c, false, false, true);
        const bool walls = !c->solid_tasks.empty() || !c->sn_tasks.empty();
        const bool side = walls && c->p2p.on;
        if (side) {
            // one-sided halo: the wall chain (ghost slopes -> solid cells -> solid halo -> solid neighbours) runs on the
            // side stream beside the cells that read no ghost data, as on a single rank
            CK(cudaEventRecord(c->ev_fork, c->stream));
            CK(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
            halo_wait(c, HK_SDF_FINAL, c->side_stream);   // (the wall kernels read the ghosts' RAW slopes)
            do_ib<D, K>(c, c->dv.df_new, c->side_stream);
            CK(cudaEventRecord(c->ev_join, c->side_stream));
        }
        for (auto& b : c->bins) if (!b.halo) launch_fused<D, K>(c, b, dt, want);
        finish_slope_halo<D, K>(c);
        if (side) CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        else do_ib<D, K>(c, c->dv.df_new);
        for (auto& b : c->bins) if (b.halo) launch_fused<D, K>(c, b, dt, want);
    }
    CK(cudaGetLastError());
    // cells that are not updated (solid ghost cells, ghosts) are refreshed in the new buffer by the IB
    // kernels / the halo exchange below
    std::swap(c->dv.df, c->dv.df_new);
    c->df_parity ^= 1;
    send_df_halo(c);
    fetch_residual<D, K>(c, want, res_out);
}

template <class F>
int guarded(kamr_ctx* c, F&& f) {
    if (!c) return 1;
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        c->err = e.what();
        return 2;
    } catch (...) {
        c->err = "unknown error";
        return 3;
    }
}

#define DISPATCH(c, fn, ...)                                                            \
    do {                                                                                \
        if ((c)->D == 2 && (c)->K == 2) fn<2, 2>(__VA_ARGS__);                          \
        else if ((c)->D == 3 && (c)->K == 1) fn<3, 1>(__VA_ARGS__);                     \
        else throw Fail("unsupported DIM/NDF (2D2F and 3D1F are built)");               \
    } while (0)

}  // namespace

// ==================================================================================================
extern "C" {

int kamr_version(void) { return KAMR_VERSION; }

const char* kamr_last_error(const kamr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int kamr_create(const kamr_config* cfg, kamr_ctx** out) {
    if (!cfg || !out) { g_create_err = "null argument"; return 1; }
    kamr_ctx* c = nullptr;
    try {
        if (!((cfg->dim == 2 && cfg->ndf == 2) || (cfg->dim == 3 && cfg->ndf == 1)))
            throw Fail("unsupported DIM/NDF: the reference's kinetics exist for 2D2F and 3D1F only (lib/KitCore)");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Fail(std::string("no CUDA device available (libkamr has no CPU fallback): ") + cudaGetErrorString(e));
        if (cfg->device < 0 || cfg->device >= ndev) throw Fail("device ordinal out of range");
        if (cfg->flux_type != KAMR_FLUX_CAIDVM && cfg->flux_type != KAMR_FLUX_DVM)
            throw Fail("unsupported flux type (CAIDVM and DVM are built; UGKS is 2-D only and broken in the reference)");
        if (cfg->marching != KAMR_MARCH_CAIDVM && cfg->marching != KAMR_MARCH_CIP && cfg->marching != KAMR_MARCH_EULER)
            throw Fail("unsupported time marching (CAIDVM_Marching, CIP_Marching and Euler are built)");
        CK(cudaSetDevice(cfg->device));
        c = new kamr_ctx();
        CK(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
        c->cfg = *cfg;
        c->D = cfg->dim; c->K = cfg->ndf; c->M = cfg->dim + 2;
        c->gas = GasPar{cfg->K, cfg->Pr, cfg->gamma, cfg->omega, cfg->mu_ref, cfg->flux_type, cfg->marching};
        if (cfg->stream) c->stream = (cudaStream_t)cfg->stream;
        else { CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
        {   // the wall kernels are short, narrow (a few hundred CTAs) and feed the last phase kernel of the step: with
            // the highest priority their CTAs are placed first and the regular cells' phase kernel fills the rest
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi));
        }
        {
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
        }
        CK(cudaEventCreateWithFlags(&c->ev_put_ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_put_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        CK(cudaMallocHost((void**)&c->h_res, 64 * sizeof(double)));
        CK(cudaMallocHost((void**)&c->h_err, sizeof(int)));
        *c->h_err = 0;
        *out = c;
        return 0;
    } catch (const std::exception& e) {
        g_create_err = e.what();
        delete c;
        return 2;
    }
}

int kamr_destroy(kamr_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    c->free_topology();
    c->mig.release();
    for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
    if (c->comm_stream) { cudaStreamSynchronize(c->comm_stream); cudaStreamDestroy(c->comm_stream); }
    if (c->ev_put_ready) cudaEventDestroy(c->ev_put_ready);
    if (c->ev_put_done) cudaEventDestroy(c->ev_put_done);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int kamr_comm_unique_id(void* id128) {
    std::string err;
    if (!nccl().load(err)) { g_create_err = err; return 2; }
    ncclUniqueId id;
    int r = nccl().GetUniqueId(&id);
    if (r) { g_create_err = nccl().GetErrorString(r); return 2; }
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int kamr_comm_init(kamr_ctx* c, const void* id128) {
    return guarded(c, [&] {
        if (c->cfg.nranks <= 1) return;
        std::string err;
        if (!nccl().load(err)) throw Fail(err);
        CK(cudaSetDevice(c->cfg.device));
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        NCK(nccl().CommInitRank(&c->comm, c->cfg.nranks, id, c->cfg.rank));
    });
}

int kamr_upload_topology(kamr_ctx* c, const kamr_mesh* m) {
    return guarded(c, [&] {
        if (!m) throw Fail("null mesh");
        CK(cudaSetDevice(c->cfg.device));
        struct Clear { kamr_ctx* c; ~Clear() { c->h_vmid = nullptr; } } clear_{c};   // the host array is the caller's
        if (!c->cells.empty()) {   // the peers' last puts into the arrays about to be freed must have landed
            halo_finish_df(c);
            halo_join_puts(c);
            CK(cudaStreamSynchronize(c->stream));
        }
        build_topology(c, m);
    });
}

int kamr_upload_state(kamr_ctx* c, const double* df, const double* w, const double* prim) {
    return guarded(c, [&] {
        CK(cudaSetDevice(c->cfg.device));
        if (c->cells.empty()) throw Fail("upload_topology first");
        halo_finish_df(c);
        halo_join_puts(c);
        c->sw_valid = false;
        if (df) copy_points(c, c->dv.df, nullptr, df, c->K, true);
        const size_t nb = (size_t)c->n_local * c->M * sizeof(double);
        if (w) CK(cudaMemcpyAsync(c->dv.w, w, nb, cudaMemcpyHostToDevice, c->stream));
        if (prim) CK(cudaMemcpyAsync(c->dv.prim, prim, nb, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    });
}

int kamr_upload_aux(kamr_ctx* c, const double* sdf, const double* flux, const double* mflux) {
    return guarded(c, [&] {
        CK(cudaSetDevice(c->cfg.device));
        if (c->cells.empty()) throw Fail("upload_topology first");
        if (sdf) {
            copy_points(c, c->dv.sdf, nullptr, sdf, c->K * c->D, true);
            DISPATCH(c, run_limit, c, c->d_limit_cells, (int)c->limit_cells.size());
            c->raw_sdf_valid = true;
            c->sw_valid = false;
        }
        if (flux) { ensure_flux(c); copy_points(c, c->dv.flux, nullptr, flux, c->K, true); }
        if (mflux) CK(cudaMemcpyAsync(c->dv.mflux, mflux, (size_t)c->n_local * c->M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    });
}

int kamr_download_state(kamr_ctx* c, uint32_t mask, double* df, double* sdf, double* flux, double* w, double* prim,
                        double* qf, double* sw, double* mflux) {
    return guarded(c, [&] {
        CK(cudaSetDevice(c->cfg.device));
        if (c->cells.empty()) throw Fail("upload_topology first");
        const int M = c->M, D = c->D;
        halo_finish_df(c);
        halo_join_puts(c);
        if ((mask & KAMR_DL_DF) && df) copy_points(c, c->dv.df, df, nullptr, c->K, false);
        if ((mask & KAMR_DL_SDF) && sdf) {
            if (!c->raw_sdf_valid)
                throw Fail("raw sdf is not resident: the fused step keeps only the limited slopes; call kamr_slope "
                           "or set KAMR_OPT_KEEP_SDF before the step whose slopes the host needs");
            copy_points(c, c->dv.sdf, sdf, nullptr, c->K * D, false);
        }
        if ((mask & KAMR_DL_FLUX) && flux) { ensure_flux(c); copy_points(c, c->dv.flux, flux, nullptr, c->K, false); }
        const size_t nl = (size_t)c->n_local;
        if ((mask & KAMR_DL_W) && w) CK(cudaMemcpyAsync(w, c->dv.w, nl * M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((mask & KAMR_DL_PRIM) && prim) CK(cudaMemcpyAsync(prim, c->dv.prim, nl * M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((mask & KAMR_DL_QF) && qf) CK(cudaMemcpyAsync(qf, c->dv.qf, nl * D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((mask & KAMR_DL_SW) && sw)   // local cells and, behind them, the ghosts (filled by kamr_slope's sw halo)
            CK(cudaMemcpyAsync(sw, c->dv.sw, (nl + (size_t)c->n_ghost) * M * D * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if ((mask & KAMR_DL_MFLUX) && mflux) CK(cudaMemcpyAsync(mflux, c->dv.mflux, nl * M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        sync_and_check(c);
    });
}

// partition migration / selective download: the listed cells' df blocks packed back to back on the device, one transfer
static void transfer_cells(kamr_ctx* c, int n, const int32_t* cells, double* df_rw, const double* df_ro, double* w_rw,
                           const double* w_ro, bool to_device) {
    if (n <= 0) return;
    if (c->cells.empty()) throw Fail("upload_topology first");
    const int K = c->K, M = c->M;
    std::vector<long long> off(n + 1, 0);
    for (int q = 0; q < n; ++q) {
        if (cells[q] < 0 || cells[q] >= c->n_local) throw Fail("cell id out of range (local cells only)");
        off[q + 1] = off[q] + c->cells[cells[q]].n;
    }
    halo_finish_df(c);
    halo_join_puts(c);
    const size_t total = (size_t)off[n] * K;
    int* d_list = nullptr; long long* d_off = nullptr; double* d_pack = nullptr; double* d_w = nullptr;
    struct Free { int*& a; long long*& b; double*& p; double*& w; ~Free() { cudaFree(a); cudaFree(b); cudaFree(p); cudaFree(w); } } fr{d_list, d_off, d_pack, d_w};
    CK(cudaMalloc((void**)&d_list, sizeof(int) * n));
    CK(cudaMalloc((void**)&d_off, sizeof(long long) * (n + 1)));
    CK(cudaMemcpyAsync(d_list, cells, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, c->stream));
    const int grid = std::min(n, 148 * 16);
    if ((df_rw || df_ro) && total) {
        CK(cudaMalloc((void**)&d_pack, sizeof(double) * total));
        if (to_device) {
            CK(cudaMemcpyAsync(d_pack, df_ro, sizeof(double) * total, cudaMemcpyHostToDevice, c->stream));
            repack_list_kernel<<<grid, 256, 0, c->stream>>>(c->dv.cells, d_list, d_off, n, K, c->dv.df, d_pack, 1);
        } else {
            repack_list_kernel<<<grid, 256, 0, c->stream>>>(c->dv.cells, d_list, d_off, n, K, c->dv.df, d_pack, 0);
            CK(cudaMemcpyAsync(df_rw, d_pack, sizeof(double) * total, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    if (w_rw || w_ro) {   // (DIM+2) doubles per cell: gathered on the host side of one small transfer
        std::vector<double> hw((size_t)c->n_local * M);
        CK(cudaMemcpyAsync(hw.data(), c->dv.w, sizeof(double) * hw.size(), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (to_device) {
            for (int q = 0; q < n; ++q) memcpy(&hw[(size_t)cells[q] * M], w_ro + (size_t)q * M, sizeof(double) * M);
            CK(cudaMemcpyAsync(c->dv.w, hw.data(), sizeof(double) * hw.size(), cudaMemcpyHostToDevice, c->stream));
        } else {
            for (int q = 0; q < n; ++q) memcpy(w_rw + (size_t)q * M, &hw[(size_t)cells[q] * M], sizeof(double) * M);
        }
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
}

int kamr_pack_cells(kamr_ctx* c, int32_t n, const int32_t* cells, double* df, double* w) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); transfer_cells(c, n, cells, df, nullptr, w, nullptr, false); });
}
int kamr_unpack_cells(kamr_ctx* c, int32_t n, const int32_t* cells, const double* df, const double* w) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); transfer_cells(c, n, cells, nullptr, df, nullptr, w, true); });
}

// Partition migration device to device.  Phase 1 on the OLD topology: the listed cells' df, w and prim are packed per
// destination rank (list order kept within a destination) and exchanged with ncclSend/ncclRecv; cells whose destination
// is this rank are copied on the device.  The arrivals are kept in a staging buffer in ascending source-rank order.
static void migrate_begin(kamr_ctx* c, int n_send, const int32_t* cells, const int32_t* dest, int n_src,
                          const int32_t* src_rank, const int32_t* src_cells, const int64_t* src_points) {
    if (c->cells.empty()) throw Fail("upload_topology first");
    if (c->mig.pending) throw Fail("kamr_migrate_begin: a migration is already pending (call kamr_migrate_finish)");
    if (n_send < 0 || n_src < 0 || (n_send && (!cells || !dest)) || (n_src && (!src_rank || !src_cells || !src_points)))
        throw Fail("kamr_migrate_begin: bad arguments");
    const int K = c->K, M = c->M, me = c->cfg.rank, nr = std::max(1, (int)c->cfg.nranks);
    // ---- sends, grouped by destination (stable)
    std::vector<int> order(n_send);
    for (int q = 0; q < n_send; ++q) {
        if (cells[q] < 0 || cells[q] >= c->n_local) throw Fail("kamr_migrate_begin: cell id outside [0, n_local)");
        if (dest[q] < 0 || dest[q] >= nr) throw Fail("kamr_migrate_begin: destination rank out of range");
        order[q] = q;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dest[a] < dest[b]; });
    std::vector<int> list(n_send);
    std::vector<long long> off(n_send + 1, 0);
    std::vector<long long> d_cells(nr, 0), d_points(nr, 0), d_cell0(nr, 0), d_point0(nr, 0);
    for (int q = 0; q < n_send; ++q) {
        const int cell = cells[order[q]], r = dest[order[q]];
        list[q] = cell;
        off[q + 1] = off[q] + c->cells[cell].n;
        if (d_cells[r] == 0) { d_cell0[r] = q; d_point0[r] = off[q]; }
        d_cells[r] += 1; d_points[r] += c->cells[cell].n;
    }
    // ---- arrivals, ascending source rank
    long long rc = 0, rp = 0;
    std::vector<long long> s_cell0(n_src), s_point0(n_src);
    bool remote = false;
    for (int i = 0; i < n_src; ++i) {
        if (src_rank[i] < 0 || src_rank[i] >= nr || (i && src_rank[i] <= src_rank[i - 1]))
            throw Fail("kamr_migrate_begin: src_rank must ascend and lie in [0, nranks)");
        if (src_cells[i] < 0 || src_points[i] < 0) throw Fail("kamr_migrate_begin: negative source counts");
        if (src_rank[i] == me && (src_cells[i] != d_cells[me] || src_points[i] != d_points[me]))
            throw Fail("kamr_migrate_begin: the counts of the cells this rank keeps do not match its own send list");
        s_cell0[i] = rc; s_point0[i] = rp;
        rc += src_cells[i]; rp += src_points[i];
        remote = remote || src_rank[i] != me;
    }
    for (int r = 0; r < nr; ++r) remote = remote || (r != me && d_cells[r] > 0);
    if (remote && !c->comm) throw Fail("kamr_migrate_begin: cells move between ranks but kamr_comm_init was not called");
    if (d_cells[me] > 0) {
        bool found = false;
        for (int i = 0; i < n_src; ++i) found = found || src_rank[i] == me;
        if (!found) throw Fail("kamr_migrate_begin: cells stay on this rank but src_rank does not list it");
    }
    halo_finish_df(c);
    halo_join_puts(c);
    double *s_df = nullptr, *s_w = nullptr, *s_prim = nullptr;
    int* d_list = nullptr; long long* d_off = nullptr;
    struct Free { double*& a; double*& b; double*& p; int*& l; long long*& o; ~Free() { cudaFree(a); cudaFree(b); cudaFree(p); cudaFree(l); cudaFree(o); } } fr{s_df, s_w, s_prim, d_list, d_off};
    auto& g = c->mig;
    try {
        if (n_send) {
            CK(cudaMalloc((void**)&s_df, sizeof(double) * std::max<size_t>(1, (size_t)off[n_send] * K)));
            CK(cudaMalloc((void**)&s_w, sizeof(double) * (size_t)n_send * M));
            CK(cudaMalloc((void**)&s_prim, sizeof(double) * (size_t)n_send * M));
            CK(cudaMalloc((void**)&d_list, sizeof(int) * n_send));
            CK(cudaMalloc((void**)&d_off, sizeof(long long) * (n_send + 1)));
            CK(cudaMemcpyAsync(d_list, list.data(), sizeof(int) * n_send, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * (n_send + 1), cudaMemcpyHostToDevice, c->stream));
            Launch L_(c, KID_PACK);
            repack_list_kernel<<<std::min(n_send, 148 * 16), 256, 0, c->stream>>>(c->dv.cells, d_list, d_off, n_send, K, c->dv.df, s_df, 0);
            const int gr = (int)std::min<long long>(((long long)n_send * M + 255) / 256, 1024);
            gather_rows_kernel<<<gr, 256, 0, c->stream>>>(d_list, n_send, M, c->dv.w, s_w);
            gather_rows_kernel<<<gr, 256, 0, c->stream>>>(d_list, n_send, M, c->dv.prim, s_prim);
            CK(cudaGetLastError());
        }
        CK(cudaMalloc((void**)&g.d_df, sizeof(double) * std::max<size_t>(1, (size_t)rp * K)));
        CK(cudaMalloc((void**)&g.d_w, sizeof(double) * std::max<size_t>(1, (size_t)rc * M)));
        CK(cudaMalloc((void**)&g.d_prim, sizeof(double) * std::max<size_t>(1, (size_t)rc * M)));
        // packed block of a cell: n*K doubles, so a destination's segment starts at point0*K
        if (remote) NCK(nccl().GroupStart());
        for (int r = 0; r < nr; ++r) {
            if (r == me || d_cells[r] == 0) continue;
            NCK(nccl().Send(s_df + d_point0[r] * K, (size_t)d_points[r] * K, ncclFloat64, r, c->comm, c->stream));
            NCK(nccl().Send(s_w + d_cell0[r] * M, (size_t)d_cells[r] * M, ncclFloat64, r, c->comm, c->stream));
            NCK(nccl().Send(s_prim + d_cell0[r] * M, (size_t)d_cells[r] * M, ncclFloat64, r, c->comm, c->stream));
        }
        for (int i = 0; i < n_src; ++i) {
            if (src_rank[i] == me || src_cells[i] == 0) continue;
            NCK(nccl().Recv(g.d_df + s_point0[i] * K, (size_t)src_points[i] * K, ncclFloat64, src_rank[i], c->comm, c->stream));
            NCK(nccl().Recv(g.d_w + s_cell0[i] * M, (size_t)src_cells[i] * M, ncclFloat64, src_rank[i], c->comm, c->stream));
            NCK(nccl().Recv(g.d_prim + s_cell0[i] * M, (size_t)src_cells[i] * M, ncclFloat64, src_rank[i], c->comm, c->stream));
        }
        if (remote) NCK(nccl().GroupEnd());
        for (int i = 0; i < n_src; ++i) {
            if (src_rank[i] != me || src_cells[i] == 0) continue;
            {   // The raw slopes stay with the cells that stay: in the reference a kept PsData keeps its VsData.sdf while
                // an arriving one starts with zeros (Partition.jl:645-652), and on meshes with non-dyadic cell sizes
                // the next sweep projects some finer neighbours' slopes of the PREVIOUS step.
                const int D = c->D;
                const int first = (int)d_cell0[me], cnt = (int)d_cells[me];
                std::vector<long long> koff(cnt + 1);
                for (int q = 0; q <= cnt; ++q) koff[q] = off[first + q] - off[first];
                long long* d_koff = nullptr;
                CK(cudaMalloc((void**)&d_koff, sizeof(long long) * (cnt + 1)));
                CK(cudaMemcpyAsync(d_koff, koff.data(), sizeof(long long) * (cnt + 1), cudaMemcpyHostToDevice, c->stream));
                CK(cudaMalloc((void**)&g.d_sdf, sizeof(double) * std::max<size_t>(1, (size_t)d_points[me] * K * D)));
                repack_list_kernel<<<std::min(cnt, 148 * 16), 256, 0, c->stream>>>(c->dv.cells, d_list + first, d_koff, cnt, K * D, c->dv.sdf, g.d_sdf, 0);
                CK(cudaGetLastError());
                CK(cudaStreamSynchronize(c->stream));   // koff is a local
                cudaFree(d_koff);
                g.kept_first = (int)s_cell0[i]; g.kept_cells = cnt; g.kept_points = d_points[me];
            }
            CK(cudaMemcpyAsync(g.d_df + s_point0[i] * K, s_df + d_point0[me] * K, sizeof(double) * (size_t)d_points[me] * K, cudaMemcpyDeviceToDevice, c->stream));
            CK(cudaMemcpyAsync(g.d_w + s_cell0[i] * M, s_w + d_cell0[me] * M, sizeof(double) * (size_t)d_cells[me] * M, cudaMemcpyDeviceToDevice, c->stream));
            CK(cudaMemcpyAsync(g.d_prim + s_cell0[i] * M, s_prim + d_cell0[me] * M, sizeof(double) * (size_t)d_cells[me] * M, cudaMemcpyDeviceToDevice, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
    } catch (...) {
        g.release();
        throw;
    }
    g.pending = true; g.points = rp; g.cells = (int)rc;
}

// Phase 2, after kamr_upload_topology of the NEW partition: arrival q (ascending source rank, the sender's list order
// within a source) becomes local cell recv_cells[q].
static void migrate_finish(kamr_ctx* c, int n_recv, const int32_t* recv_cells) {
    auto& g = c->mig;
    if (!g.pending) throw Fail("kamr_migrate_finish: no migration pending");
    if (c->cells.empty()) throw Fail("upload_topology first");
    struct Release { kamr_ctx::Migrate& g; ~Release() { g.release(); } } rel{g};
    if (n_recv != g.cells) throw Fail("kamr_migrate_finish: n_recv differs from the number of cells that arrived");
    if (n_recv == 0) return;
    if (!recv_cells) throw Fail("kamr_migrate_finish: recv_cells is NULL");
    const int K = c->K, M = c->M;
    std::vector<long long> off(n_recv + 1, 0);
    for (int q = 0; q < n_recv; ++q) {
        if (recv_cells[q] < 0 || recv_cells[q] >= c->n_local) throw Fail("kamr_migrate_finish: cell id outside [0, n_local)");
        off[q + 1] = off[q] + c->cells[recv_cells[q]].n;
    }
    if (off[n_recv] != g.points)
        throw Fail("kamr_migrate_finish: the velocity grids of recv_cells do not add up to the points that arrived");
    halo_finish_df(c);
    halo_join_puts(c);
    int* d_list = nullptr; long long* d_off = nullptr;
    struct Free { int*& l; long long*& o; ~Free() { cudaFree(l); cudaFree(o); } } fr{d_list, d_off};
    CK(cudaMalloc((void**)&d_list, sizeof(int) * n_recv));
    CK(cudaMalloc((void**)&d_off, sizeof(long long) * (n_recv + 1)));
    CK(cudaMemcpyAsync(d_list, recv_cells, sizeof(int) * n_recv, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * (n_recv + 1), cudaMemcpyHostToDevice, c->stream));
    {
        Launch L_(c, KID_UNPACK);
        repack_list_kernel<<<std::min(n_recv, 148 * 16), 256, 0, c->stream>>>(c->dv.cells, d_list, d_off, n_recv, K, c->dv.df, g.d_df, 1);
        const int gr = (int)std::min<long long>(((long long)n_recv * M + 255) / 256, 1024);
        scatter_rows_kernel<<<gr, 256, 0, c->stream>>>(d_list, n_recv, M, g.d_w, c->dv.w);
        scatter_rows_kernel<<<gr, 256, 0, c->stream>>>(d_list, n_recv, M, g.d_prim, c->dv.prim);
    }
    CK(cudaGetLastError());
    if (g.kept_cells > 0 && g.d_sdf) {   // the kept cells' raw slopes (what their VsData.sdf would still hold)
        const int D = c->D, first = g.kept_first, cnt = g.kept_cells;
        std::vector<long long> koff(cnt + 1);
        for (int q = 0; q <= cnt; ++q) koff[q] = off[first + q] - off[first];
        if (koff[cnt] != g.kept_points) throw Fail("kamr_migrate_finish: the kept cells' velocity grids changed size");
        long long* d_koff = nullptr;
        CK(cudaMalloc((void**)&d_koff, sizeof(long long) * (cnt + 1)));
        CK(cudaMemcpyAsync(d_koff, koff.data(), sizeof(long long) * (cnt + 1), cudaMemcpyHostToDevice, c->stream));
        repack_list_kernel<<<std::min(cnt, 148 * 16), 256, 0, c->stream>>>(c->dv.cells, d_list + first, d_koff, cnt, K * D, c->dv.sdf, g.d_sdf, 1);
        cudaError_t e = cudaGetLastError();
        cudaStreamSynchronize(c->stream);
        cudaFree(d_koff);
        CK(e);
    }
    CK(cudaStreamSynchronize(c->stream));
    c->sw_valid = false; c->raw_sdf_valid = false;
}

int kamr_migrate_begin(kamr_ctx* c, int32_t n_send, const int32_t* cells, const int32_t* dest_rank, int32_t n_src,
                       const int32_t* src_rank, const int32_t* src_cells, const int64_t* src_points) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); migrate_begin(c, n_send, cells, dest_rank, n_src, src_rank, src_cells, src_points); });
}
int kamr_migrate_finish(kamr_ctx* c, int32_t n_recv, const int32_t* recv_cells) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); migrate_finish(c, n_recv, recv_cells); });
}

int kamr_slope(kamr_ctx* c) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_slope, c, true, true); });
}
int kamr_ps_criterion(kamr_ctx* c, double threshold, double* lohner_out, double* sensor_out) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_ps_criterion, c, threshold, lohner_out, sensor_out); });
}
int kamr_vs_resolution(kamr_ctx* c, const kamr_vs_adapt* par, double* out) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_vs_resolution, c, par, out); });
}
int kamr_vs_criterion(kamr_ctx* c, const kamr_vs_adapt* par, uint8_t* refine_flag, uint8_t* coarsen_ok) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_vs_criterion, c, par, refine_flag, coarsen_ok); });
}
int kamr_project_cells(kamr_ctx* c, int32_t n, const int32_t* cells) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_project_cells, c, (int)n, cells); });
}
int kamr_flux(kamr_ctx* c, double dt) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_flux, c, dt); });
}
int kamr_iterate(kamr_ctx* c, double dt, int32_t want_residual, double* res_out) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_iterate, c, dt, want_residual, res_out); });
}
int kamr_step(kamr_ctx* c, double dt, int32_t want_residual, double* res_out) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); DISPATCH(c, do_step, c, dt, want_residual, res_out); });
}
int kamr_exchange_df(kamr_ctx* c) {
    return guarded(c, [&] {
        CK(cudaSetDevice(c->cfg.device));
        send_df_halo(c);
        halo_finish_df(c);
    });
}
int kamr_sync(kamr_ctx* c) {
    return guarded(c, [&] { CK(cudaSetDevice(c->cfg.device)); halo_finish_df(c); halo_join_puts(c); sync_and_check(c); });
}

int kamr_get_stats(kamr_ctx* c, kamr_stats* out) {
    return guarded(c, [&] {
        memset(out, 0, sizeof(*out));
        out->n_phase_local = c->n_phase_local;
        out->n_points_total = c->npts_host;
        out->n_relations = (long long)c->rel_off.size();
        out->n_slots = (long long)c->slots.size();
        out->kernel_launches = c->launches;
        out->device_bytes = c->device_bytes;
        out->halo_bytes_per_step = c->halo_bytes_step;
        out->n_levels = (int)c->slope_stages.size();
        out->fused_cells = c->fused_cells;
    });
}

int kamr_set_option(kamr_ctx* c, int32_t option, int32_t value) {
    return guarded(c, [&] {
        if (option == KAMR_OPT_KEEP_SDF) c->keep_sdf = value != 0;
        else throw Fail("unknown option");
    });
}

int kamr_profile_enable(kamr_ctx* c, int32_t on) {
    return guarded(c, [&] {
        CK(cudaSetDevice(c->cfg.device));
        CK(cudaStreamSynchronize(c->stream));
        for (auto& r : c->prof_recs) { c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b); }
        c->prof_recs.clear();
        c->profiling = on != 0;
    });
}

int kamr_profile_read(kamr_ctx* c, kamr_kernel_time* out, int32_t cap, int32_t* n) {
    return guarded(c, [&] {
        CK(cudaSetDevice(c->cfg.device));
        CK(cudaStreamSynchronize(c->stream));
        double ms[KID_COUNT] = {0};
        long long cnt[KID_COUNT] = {0};
        for (auto& r : c->prof_recs) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, r.a, r.b));
            ms[r.kid] += t; cnt[r.kid]++;
            c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
        }
        c->prof_recs.clear();
        int k = 0;
        for (int q = 0; q < KID_COUNT; ++q) {
            if (!cnt[q]) continue;
            if (k >= cap) throw Fail("buffer too small");
            memset(&out[k], 0, sizeof(out[k]));
            strncpy(out[k].name, kKernelNames[q], sizeof(out[k].name) - 1);
            out[k].launches = cnt[q]; out[k].total_ms = ms[q];
            ++k;
        }
        *n = k;
    });
}

int kamr_debug_exp_nonpos(kamr_ctx* c, const double* x, double* y, int64_t n) {
    return guarded(c, [&] {
        CK(cudaSetDevice(c->cfg.device));
        if (n <= 0) return;
        double *dx = nullptr, *dy = nullptr;
        CK(cudaMalloc((void**)&dx, n * sizeof(double)));
        CK(cudaMalloc((void**)&dy, n * sizeof(double)));
        CK(cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        exp_nonpos_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, c->stream>>>(dx, dy, n);
        CK(cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        cudaError_t e = cudaStreamSynchronize(c->stream);
        cudaFree(dx); cudaFree(dy);
        CK(e);
    });
}

int kamr_get_pair_map(kamr_ctx* c, int32_t ga, int32_t gb, int32_t* start, int32_t cap) {
    int rc = 0;
    int g = guarded(c, [&] {
        if (ga < 0 || gb < 0 || ga >= c->n_grid || gb >= c->n_grid) throw Fail("grid id out of range");
        if (c->grid_canon[ga] == c->grid_canon[gb]) { rc = 1; return; }
        auto it = c->rel_id.find(std::make_pair(c->grid_canon[ga], c->grid_canon[gb]));
        if (it == c->rel_id.end()) throw Fail("no relation between these grids in the current topology");
        const int na = c->grid_n[ga];
        if (cap < na + 1) throw Fail("buffer too small");
        memcpy(start, c->pm_start.data() + c->rel_off[it->second], sizeof(int) * (na + 1));
    });
    return g ? g + 1 : rc;  // 0 ok, 1 identity, >=2 error
}

int kamr_get_cell_slots(kamr_ctx* c, int32_t cell, int32_t* face, int32_t* sign, int32_t cap, int32_t* n) {
    return guarded(c, [&] {
        if (cell < 0 || cell >= c->n_local) throw Fail("cell out of range");
        const CellInfo& ci = c->cells[cell];
        const int ns = ci.slot_end - ci.slot_begin;
        if (cap < ns) throw Fail("buffer too small");
        for (int q = 0; q < ns; ++q) {
            face[q] = c->slots[ci.slot_begin + q].face;
            sign[q] = c->slots[ci.slot_begin + q].is_here ? 1 : -1;
        }
        *n = ns;
    });
}

}  // extern "C"
