// kamr_types.h — device-side data layout of libkamr.
#pragma once
#include <stdint.h>

namespace kamr {

constexpr int MAXD = 3;
constexpr int MAXM = 5;
// face slots of one cell: 2*DIM sides x 2^(DIM-1) sub-faces
template <int DIM> struct MaxSlots { static constexpr int value = 2 * DIM * (1 << (DIM - 1)); };
inline int max_slots(int dim) { return 2 * dim * (1 << (dim - 1)); }
constexpr int MAX_SLOTS = MaxSlots<MAXD>::value;
constexpr int MAX_SLOPE_NB = 24;  // DIM dirs x 2 sides x 2^(DIM-1) neighbours
constexpr int PAD = 4;          // planes padded to 4 doubles (32 B) so 128-bit loads stay aligned

constexpr int VPK_BITS = 9;       // <= 512 distinct coordinates per axis
constexpr int VPK_LEVELS = 16;    // velocity refinement levels 0..15

constexpr double EPS_KIT = 1e-12;                   // src/Abstract/Types.jl:3
constexpr double EPS_MACH = 2.220446049250313e-16;  // Julia eps(), Flux/CAIDVM.jl:137

// One physical cell (local, ghost or solid-neighbour slot).
struct CellInfo {
    long long doff;   // padded point offset of the cell's df/sdf/flux blocks
    long long goff;   // padded point offset of its velocity grid statics
    int n, np;        // velocity points, padded plane stride
    int grid;
    int bound_enc;
    int slot_begin, slot_end;
    int ps_level;
    int flags;                     // CELL_* bits
    int hot_begin;                 // first FaceRec of the cell; records sorted by (direction, side)
    int rare_begin, rare_count;    // cell-relative indices of the slots pass B visits
    unsigned char side_begin[8];   // [2*d + side] .. [2*d + side + 1]: FaceRec range of that side; [2*DIM]: total
    int sn_begin, sn_count;        // donor cells: their SolidNeighbor faces in DevView::donor_sn (positivity_preserving_ib!)
    int pad_;
    double ds[MAXD], mid[MAXD], vol;
};
// A SolidNeighbor face as its donor cell sees it in positivity_preserving_ib! (Boundary/Positivity.jl:11-27)
struct DonorSn {
    long long doff;      // the SolidNeighbor pseudo-cell's block (its vs_data.flux / .sdf live there)
    int dir, pad_;
    double rot;          // get_rot(faceid)
    double area;         // rot * prod_{t != dir} ds_t of the donor
    double fmid[MAXD];   // face midpoint: donor midpoint shifted by -rot ds_dir / 2 along dir
    double snmid[MAXD];  // SolidNeighbor.midpoint (= its solid cell's)
};
enum CellFlags : int {
    CELL_HAS_MAXWELL_WALL = 1,
    CELL_HAS_MAPPED = 4,   // some fluid/fluid face leads to a neighbour on a different velocity grid
    CELL_REGULAR_MAPPED = 8,  // regular topology and geometry, but some neighbour is on a different velocity grid
    CELL_REGULAR = 2,  // each of the 2*DIM sides is one fluid/fluid face to a neighbour on the same velocity grid
};

enum SlotKind : int {
    SLOT_INNER = 0,       // fluid/fluid: limited reconstruction on both sides (CAIDVM.jl:113-114)
    SLOT_NBR_SOLID = 1,   // there is a solid cell / SolidNeighbor (CAIDVM.jl:110-111)
    SLOT_BC_MAXWELL = 2,  // CAIDVM.jl:4
    SLOT_BC_INFLOW = 3,   // CAIDVM.jl:29
    SLOT_BC_UNIFORM = 4,  // CAIDVM.jl:53
    SLOT_BC_INTERP = 5,   // CAIDVM.jl:71
};

// One (cell, face) incidence: what a cell gathers through one face.  Everything the kernel needs
// about the neighbour is resolved at flatten time so that the device never chases cell records.
struct Slot {
    long long nbr_doff;  // neighbour's point offset (df/sdf blocks)
    long long nbr_goff;  // neighbour's velocity-grid offset
    long long rel_off;   // offset of the pair map (own grid -> nbr grid) in pm_start; -1: identical grids
    int nbr;             // neighbour cell id, -1 for domain faces
    int nbr_np;
    int dir;
    int kind;            // SlotKind
    int is_here;         // 1: this cell is the face's here side, 0: there side
    int face;            // index in the host face list
    double rot;          // face rot (+1/-1)
    double area;         // signed: +rot*A for here, -rot*A for there (Flux.jl:84-136)
    double fmid[MAXD];     // face midpoint
    double own_mid[MAXD];  // midpoint this cell carries in the face record (periodic aliases are shifted)
    double nbr_mid[MAXD];  // neighbour midpoint of the face record
    double nds[MAXD];    // neighbour ds
    double bc[MAXM];
};

// A fluid/fluid face as the flux gather sees it from one cell (built at flatten time from Slot; pass A of the phase
// kernel reads nothing else).  side 0: the face at the low end of the direction, side 1: the high end.  Through the
// low face the cell is upwind for points with v_d < 0 and the neighbour for v_d > 0; the high face the other way round.
struct FaceRec {
    long long nf_off;    // neighbour df block offset (doubles)
    long long nsl_off;   // neighbour limited-slope block offset (doubles)
    int np;              // neighbour plane stride
    int flags;           // bit1: identical velocity grids (plain gather); clear: pair-mapped gather via rel_off
    long long rel_off;   // pair map own grid -> neighbour grid in pm_start (flags bit1 clear)
    long long ngoff;     // neighbour's velocity-grid statics offset
    double area;         // signed, as Slot::area
    double fmid[MAXD], own_mid[MAXD], nbr_mid[MAXD];
};

// A REGULAR cell as phase_regular_kernel sees it: one record per CTA.  Regular: every side is one fluid/fluid face to a
// same-size neighbour on the same velocity grid, and in the coordinates transverse to a face the face midpoint and the
// neighbour midpoint equal the cell's own bit for bit (build_topology checks), so only the normal coordinate of a
// side is carried.
struct RegSide {
    long long ndoff;     // neighbour's padded point offset (df block at ndoff*NDF, slope block at ndoff*NDF*DIM)
    double area;         // signed, as Slot::area
    double fmid;         // face midpoint, normal coordinate
    double nmid;         // neighbour midpoint, normal coordinate
    long long rel_off;   // pair map own grid -> neighbour grid in pm_start; -1: identical grids (always, in plain bins)
    long long ngoff;     // neighbour's velocity-grid statics offset
    int np, pad_;        // neighbour's plane stride
};
struct RegCell {
    long long doff, goff;
    int n, np, cell, pad_;
    double vol, mid[MAXD];
    RegSide side[2 * MAXD];   // [2*d + side]
};

// Slope stencil of one (cell, direction): Flux/Slope.jl:458-771, 849-945 resolved at flatten time.
struct SlopeNbr {
    long long doff;      // neighbour's point offset
    long long goff;      // neighbour's velocity-grid offset
    long long rel_off;   // pair map offset, -1 identity
    int np;
    int proj;            // transverse projection active for this neighbour
    double dm[MAXD];     // own midpoint - neighbour midpoint (0 along the slope direction)
};
enum SlopeMode : int { SLOPE_ZERO = 0, SLOPE_BOUND = 1, SLOPE_INNER = 2, SLOPE_KEEP = 3 };
struct SlopeDir {
    int mode;            // SlopeMode; BOUND uses side A only
    int nA, nB;          // neighbours on side A / B: entries [nb_begin, +nA) and [nb_begin+nA, +nB)
    int nb_begin;        // into the SlopeNbr array
    double invA, invB;   // 1 / (divisor * number of neighbours), signed
};
struct SlopeTask {
    int cell;
    int flags;           // bit0: write raw sdf (somebody reads it)
    int dep_begin, dep_count;  // cells (in DevView::slope_deps) whose finished slopes this task projects, when they
                               // are computed by the same launch (single-launch dependency sweep)
    SlopeDir d[MAXD];
};
// A REGULAR slope stencil: in every direction one same-level fluid neighbour per side on the same velocity grid,
// no transverse projection (the bulk of every mesh).  Everything the kernel needs in one record.
struct SlopeReg {
    long long doff;
    int n, np;
    int flags, pad_;     // bit0: write raw sdf
    double ds[MAXD];
    long long nb_doff[2 * MAXD];   // [2*d + side]
    double inv[2 * MAXD];          // 1/dsL, 1/dsR (signed) as SlopeDir::invA/invB
};
// The same stencil with neighbours on other velocity grids (pair-mapped gather): slope_regular_kernel<MAPPED>
struct SlopeRegMap {
    SlopeReg r;
    long long goff;                // own velocity-grid statics (levels)
    long long nb_rel[2 * MAXD];    // pair map own grid -> neighbour grid, -1 identical
    long long nb_goff[2 * MAXD];
    int nb_np[2 * MAXD];
};

// Immersed boundary (kernel d): tables resolved at flatten time from kamr_ib.
struct IbNbr {            // one fluid cell a solid cell / an image point extrapolates from
    long long doff, goff; // its df/sdf block and velocity-grid statics
    long long rel_off;    // pair map (target grid -> this cell's grid), -1 identical
    int np, pad_;
    double mid[MAXD];
};
struct SolidTask {        // update_solid_cell!, Boundary/Immersed_boundary.jl:122-141
    int cell;
    int nb_begin, nb_count;
    int pad_;
};
struct SnTask {           // update_solid_neighbor!, Boundary/Immersed_boundary.jl:436-480
    int sn_cell, donor, solid, dir;
    int nb_begin, nb_count;        // donor's fluid neighbours followed by the donor itself
    int cvc_begin, cvc_count;      // cut velocity cells (sorted by point index)
    long long rel_ps;              // pair map donor grid -> solid cell's grid, -1 identical
    double aux[MAXD], normal[MAXD], bc[MAXM];
};

// Segment copy descriptor (halo pack / unpack)
struct CopySeg {
    long long src, dst;   // offsets in doubles
    long long len;
};

// One-sided halo over NVLink: a mirror cell's block is stored straight into the peer's ghost block
// (the peer's arrays are mapped through CUDA IPC at re-flatten time).
struct PutSeg {
    long long src, dst;   // offsets in doubles in the local / the peer's array
    int len;              // doubles
    int peer;             // index into the per-peer base-pointer table
};

// Device pointers handed to the kernels.
struct DevView {
    const CellInfo* cells;
    const Slot* slots;
    const FaceRec* hot;
    const int* rare;
    const SlopeNbr* slope_nb;
    const int* slope_deps;
    int* slope_done;            // per cell: epoch of the last finished slope task (dependency flags)
    int* err_flag;              // raised by a kernel that gave up waiting (slope dependency sweep); checked at sync points
    const int8_t* v_level;
    const unsigned char* v_sign; // per point: bit d = (v_d > 0)
    const double* v_weight;
    const double* v_mid;
    // Packed velocity-grid statics: the coordinates of all velocity points of a run take few distinct
    // values per axis (root index x refinement offsets), and a point's weight is a function of its level alone
    // (Velocity_space/Rebuild.jl:60), so one 32-bit word per point — an index into a per-axis table of the actual
    // doubles plus the level — replaces DIM+1 doubles.  The hot kernels stage the tables in shared memory and read
    // 4 bytes per point and pass instead of 8*(DIM+1); values are bit-identical to v_mid / v_weight by construction.
    const unsigned* v_pack;     // per point: axis d index in bits [VPK_BITS*d, +VPK_BITS), level in bits [27, 31)
    const double* v_tab;        // [DIM][n_vtab] coordinate tables, then [VPK_LEVELS] weights by level
    int n_vtab;
    const int* pm_start;        // concatenated pair maps
    const IbNbr* ib_nb;
    const DonorSn* donor_sn;
    const int* cvc_index;       // cut velocity cells of all SolidNeighbors
    const double* cvc_gas_w;
    const double* cvc_solid_w;
    double* df;                 // current distribution (read side of a fused step)
    double* df_new;             // write side of a fused step
    double* sdf;                // raw slopes (reference semantics)
    double* sdl;                // limited slopes r*sdf (device-internal)
    double* flux;
    double* w;
    double* prim;
    double* mflux;
    double* qf;
    double* sw;
    double* res_cell;           // [n_local * 2*(DIM+2)] residual contributions
    int n_local;
};

struct GasPar {
    double K, Pr, gamma, omega, mu_ref;
    int flux_type, marching;
};

}  // namespace kamr
