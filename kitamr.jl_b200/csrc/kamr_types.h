// kamr_types.h — device-side data layout of libkamr.
#pragma once
#include <stdint.h>

namespace kamr {

constexpr int MAXD = 3;
constexpr int MAXM = 5;
constexpr int MAX_SLOTS = 32;   // 2*DIM sides x 2^(DIM-1) sub-faces, plus slack
constexpr int PAD = 4;          // planes padded to 4 doubles (32 B) so 128-bit loads stay aligned

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
    int ps_level, pad_;
    double ds[MAXD], mid[MAXD], vol;
};

enum SlotKind : int {
    SLOT_INNER = 0,       // fluid/fluid: limited reconstruction on both sides (CAIDVM.jl:113-114)
    SLOT_NBR_SOLID = 1,   // there is a solid cell / SolidNeighbor (CAIDVM.jl:110-111)
    SLOT_BC_MAXWELL = 2,  // CAIDVM.jl:4
    SLOT_BC_INFLOW = 3,   // CAIDVM.jl:29
    SLOT_BC_UNIFORM = 4,  // CAIDVM.jl:53
    SLOT_BC_INTERP = 5,   // CAIDVM.jl:71
};

// One (cell, face) incidence: what a cell gathers through one face.
struct Slot {
    int nbr;        // neighbour cell id, -1 for domain faces
    int rel;        // pair-map id of (own grid -> nbr grid), -1 when the grids are identical
    int dir;
    int kind;       // SlotKind
    int is_here;    // 1: this cell is the face's here side, 0: there side
    int face;       // index in the host face list
    double rot;     // face rot (+1/-1)
    double area;    // signed: +rot*A for here, -rot*A for there (Flux.jl:84-136)
    double fmid[MAXD];
    double own_mid[MAXD];  // midpoint this cell carries in the face record (periodic aliases are shifted)
    double nbr_mid[MAXD];
    double bc[MAXM];
};

// Slope stencil of one (cell, direction): Flux/Slope.jl:458-771, 849-945 resolved at flatten time.
struct SlopeSide {
    int n;             // neighbours on this side (1, or 2^(DIM-1) finer cells)
    int nbr[4];
    int rel[4];        // pair-map id per neighbour (-1 identity)
    int proj[4];       // transverse projection active for this neighbour
    int pad_;
    double ds;         // divisor (signed)
    double dm[4][MAXD];
};
enum SlopeMode : int { SLOPE_ZERO = 0, SLOPE_BOUND = 1, SLOPE_INNER = 2, SLOPE_KEEP = 3 };
struct SlopeDir {
    int mode;          // SlopeMode; BOUND uses side A only
    int pad_;
    SlopeSide A, B;
};
struct SlopeTask {
    int cell;
    int pad_;
    SlopeDir d[MAXD];
};

// Segment copy descriptor (halo pack / unpack)
struct CopySeg {
    long long src, dst;   // offsets in doubles
    long long len;
};

// Device pointers handed to the kernels.
struct DevView {
    const CellInfo* cells;
    const Slot* slots;
    const int8_t* v_level;
    const double* v_weight;
    const double* v_mid;
    const int* pm_start;        // concatenated pair maps
    const long long* rel_off;   // [n_rel] offset of each map in pm_start
    double* df;                 // current distribution (read side of a fused step)
    double* df_new;             // write side of a fused step
    double* sdf;
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
