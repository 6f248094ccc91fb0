/*
 * kamr.h — C-ABI of libkamr, the B200-native replacement for KitAMR.jl's
 * per-step phase-space path   slope! -> flux! -> iterate!
 *
 * Reference seam (there is no FFI in the reference; these are the Julia entry
 * points a host shim replaces):
 *   slope!(p4est, ka)     src/Flux/Slope.jl:1047      -> kamr_slope
 *   flux!(p4est, ka)      src/Flux/Flux.jl:458        -> kamr_flux
 *   iterate!(p4est, ka)   src/Theory/Iterate.jl:5     -> kamr_iterate
 *   solve! loop body      src/Solver/Solver.jl:59-67  -> kamr_step (fused)
 *   amr_recover!          src/Solver/AMR.jl:54        -> kamr_upload_topology (re-flatten trigger)
 *   data_exchange! etc.   src/Parallel/Ghost.jl:841,896,867 -> internal NCCL halo (kamr_comm_init)
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on failure; the message
 *     is available from kamr_last_error(ctx) (or kamr_last_error(NULL) for
 *     failures of kamr_create).  No C++ exception crosses this boundary and the
 *     library never calls back into the host.
 *   - all pointers are HOST pointers owned by the caller, read (or written, for
 *     downloads) only during the call.  The library owns all device memory.
 *   - one context per rank == per GPU; calls on a context are serialised by the
 *     caller (KitAMR runs one Julia task per MPI rank).
 *   - indices are 0-based int32 unless stated; reals are fp64.
 *
 * Host array layouts (identical to the reference's Julia column-major blocks)
 *   cell ids: [0,n_local) local cells that own velocity data (fluid, donor and
 *             solid ghost cells — InsideSolidData placeholders are not listed),
 *             [n_local, n_local+n_ghost) p4est ghost cells (contiguous per
 *             source rank, Parallel/Ghost.jl:221-239),
 *             [n_local+n_ghost, n_cell) SolidNeighbor pseudo-cells
 *             (Physical_space/Types.jl:28-48).
 *   per-point state of cell c with n = n(c) points, starting at
 *   P = vs_off[c] = sum_{c'<c} n(c'):
 *      df  [ (P*NDF)      + k*n + i ]            == VsData.df[i,k]
 *      sdf [ (P*NDF*DIM)  + (d*NDF + k)*n + i ]  == VsData.sdf[i,k,d]
 *      flux[ (P*NDF)      + k*n + i ]            == VsData.flux[i,k]
 *   per-cell:  w,prim,mflux [c*(DIM+2)+m];  qf [c*DIM+d];  sw [c*(DIM+2)*DIM + d*(DIM+2)+m]
 *              (downloads cover the local cells; sw also the ghost cells behind them: kamr_slope ships the mirrors'
 *               macro slopes as sw_exchange! does, Parallel/Ghost.jl:867, for the host's Löhner sensor)
 *   velocity grids are shared: cell c uses grid g = cell_grid[c]; grid g has
 *   n_g = grid_off[g+1]-grid_off[g] points at Q = grid_off[g]:
 *      v_level [Q+i]  (Int8, VsData.level)     v_weight[Q+i] (VsData.weight)
 *      v_mid   [Q*DIM + d*n_g + i]             == VsData.midpoint[i,d]
 *   (the host shim may give every cell its own grid; sharing identical grids
 *    lets the device keep one copy — see DESIGN.md "static dedup").
 */
#ifndef KAMR_H
#define KAMR_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KAMR_VERSION 1

/* flux scheme: Solver.flux (src/Solver/Types.jl:69) */
enum { KAMR_FLUX_CAIDVM = 0, KAMR_FLUX_DVM = 1 };
/* time marching: Solver.time_marching (src/Solver/Types.jl:71).
 *   CAIDVM  iterate!(::Type{CAIDVM_Marching}), Theory/Iterate.jl:96   (kamr_step fuses flux! + iterate!)
 *   CIP     iterate!(::Type{CIP_Marching}),    Theory/I-projection.jl:161 (Newton I-projection per cell; on donor cells
 *           of an immersed boundary preceded by positivity_preserving_ib!, Boundary/Positivity.jl:1)
 *   EULER   iterate!(::Type{Euler}),           Theory/Iterate.jl:131 */
enum { KAMR_MARCH_CAIDVM = 0, KAMR_MARCH_CIP = 1, KAMR_MARCH_EULER = 2 };
/* face kinds (src/Physical_space/Types.jl:157-219, src/Boundary/Types.jl:27-33).
 * Hanging / BackHanging faces are passed as one record per (here, there) pair,
 * i.e. one per FluxData (src/Flux/Flux.jl:37-59). */
enum { KAMR_FACE_DOMAIN = 0, KAMR_FACE_FULL = 1, KAMR_FACE_HANGING = 2, KAMR_FACE_BACKHANGING = 3 };
/* domain boundary conditions (src/Flux/CAIDVM.jl:4,29,53,71) */
enum { KAMR_BC_MAXWELLIAN = 0, KAMR_BC_SUPERSONIC_INFLOW = 1, KAMR_BC_UNIFORM_OUTFLOW = 2,
       KAMR_BC_INTERPOLATED_OUTFLOW = 3 };
/* download mask bits */
enum { KAMR_DL_DF = 1, KAMR_DL_SDF = 2, KAMR_DL_FLUX = 4, KAMR_DL_W = 8, KAMR_DL_PRIM = 16,
       KAMR_DL_QF = 32, KAMR_DL_SW = 64, KAMR_DL_MFLUX = 128 };

typedef struct kamr_ctx kamr_ctx;

typedef struct kamr_config {
    int32_t dim;          /* 2 or 3 */
    int32_t ndf;          /* 2 (2D2F) or 1 (3D1F) */
    int32_t flux_type;    /* KAMR_FLUX_* */
    int32_t marching;     /* KAMR_MARCH_* */
    double  K, Pr, gamma, omega, mu_ref;  /* Gas (src/Gas/Types.jl:5-24) */
    int32_t device;       /* CUDA device ordinal */
    int32_t rank, nranks; /* MPI rank / size of the owning process */
    void*   stream;       /* optional cudaStream_t to launch on (NULL: library creates one) */
} kamr_config;

typedef struct kamr_mesh {
    /* ---- cells ---- */
    int32_t n_local, n_ghost, n_solidnbr;
    const double*  ds;        /* [n_cell*DIM]  PsData.ds */
    const double*  mid;       /* [n_cell*DIM]  PsData.midpoint */
    const int32_t* bound_enc; /* [n_cell] 0 fluid, >0 donor, <0 solid (PsData.bound_enc) */
    const int32_t* ps_level;  /* [n_cell] physical refinement level (cell_level, Slope.jl:825) */
    const int32_t* cell_grid; /* [n_cell] velocity-grid id */
    /* ---- velocity grids ---- */
    int32_t n_grid;
    const int64_t* grid_off;  /* [n_grid+1] */
    const int8_t*  v_level;
    const double*  v_weight;
    const double*  v_mid;
    /* ---- slope neighbours of local cells: PsData.neighbor (Mesh/Neighbor.jl:31-63) ----
     * entry e = c*2*DIM + (faceid-1); state: 0 boundary, 1 same, -1 coarser, 2^(DIM-1) finer */
    const int32_t* nb_state;  /* [n_local*2*DIM] */
    const int32_t* nb_off;    /* [n_local*2*DIM+1] */
    const int32_t* nb_ids;    /* neighbour cell ids (local, ghost or solid-neighbour slots) */
    int32_t ps_maxlevel;      /* Solver.AMR_PS_MAXLEVEL */
    int32_t ps_minlevel;      /* min_cell_level(ka) over ALL ranks (Slope.jl:953-967) */
    /* ---- faces: ka.kdata.field.faces, in the host's order ---- */
    int32_t n_face;
    const int32_t* face_kind;   /* KAMR_FACE_* */
    const int32_t* face_here;   /* local cell id */
    const int32_t* face_there;  /* cell id, or bc index for domain faces */
    const int32_t* face_dir;    /* 0-based direction */
    const double*  face_rot;    /* +1 / -1  (get_rot, Theory/Math.jl:2) */
    const double*  face_mid;    /* [n_face*DIM] face midpoint */
    const double*  face_there_mid; /* [n_face*DIM] there_data.midpoint (shifted for periodic aliases) */
    /* ---- domain boundary conditions ---- */
    int32_t n_bc;
    const int32_t* bc_type;     /* KAMR_BC_* */
    const double*  bc_prim;     /* [n_bc*(DIM+2)] */
    /* ---- halo (p4est ghost layer, Parallel/Ghost.jl:133-145) ---- */
    int32_t n_peer;
    const int32_t* peer_rank;   /* [n_peer] */
    const int32_t* send_off;    /* [n_peer+1] into send_cells */
    const int32_t* send_cells;  /* mirror cells (local ids) in mirror_proc_mirrors order */
    const int32_t* recv_off;    /* [n_peer+1] ghost index ranges (relative to n_local) */
    /* ---- immersed boundary (Boundary/Immersed_boundary.jl) ---- */
    const struct kamr_ib* ib;   /* NULL when no immersed boundary */
} kamr_mesh;

/* Immersed-boundary tables, built by the host at re-flatten time
 * (initialize_solid_neighbor! :282, initialize_cutted_velocity_cell :226). */
typedef struct kamr_ib {
    /* solid ghost cells (bound_enc<0): fluid neighbours used by update_solid_cell! :122 */
    int32_t n_solid;
    const int32_t* solid_cell;    /* [n_solid] local cell id */
    const int32_t* solid_nb_off;  /* [n_solid+1] */
    const int32_t* solid_nb_ids;  /* fluid neighbour cell ids (first element of each face/corner list) */
    /* solid neighbours (one per donor cell x solid face) */
    int32_t n_sn;                 /* == n_solidnbr */
    const int32_t* sn_donor;      /* [n_sn] donor cell id */
    const int32_t* sn_solid;      /* [n_sn] solid cell id (local or ghost) */
    const int32_t* sn_faceid;     /* [n_sn] 0-based face id of the donor */
    const double*  sn_aux;        /* [n_sn*DIM] wall intersection point */
    const double*  sn_normal;     /* [n_sn*DIM] */
    const double*  sn_bc;         /* [n_sn*(DIM+2)] wall prim evaluated at aux point */
    const int32_t* sn_nb_off;     /* [n_sn+1] donor's fluid neighbours used by image_df :376 */
    const int32_t* sn_nb_ids;
    /* cut velocity cells (CuttedVelocityCells, Velocity_space/Types.jl:81-94) */
    const int32_t* cvc_off;       /* [n_sn+1] */
    const int32_t* cvc_index;     /* velocity point index within the donor grid */
    const double*  cvc_gas_w;     /* gas-side weight */
    const double*  cvc_solid_w;   /* solid-side weight */
} kamr_ib;

/* lifecycle */
int  kamr_create(const kamr_config* cfg, kamr_ctx** out);
int  kamr_destroy(kamr_ctx* ctx);
const char* kamr_last_error(const kamr_ctx* ctx);
int  kamr_version(void);

/* multi-GPU: nccl_unique_id points at the 128-byte ncclUniqueId obtained on
 * rank 0 via kamr_comm_unique_id and distributed by the host (MPI_Bcast in the
 * Julia shim).  Not needed for nranks == 1.  NCCL carries the set-up handshakes; the per-step halo is one-sided
 * (mirror blocks stored into the peer's ghost blocks over NVLink through CUDA IPC mappings), or
 * two-sided ncclSend/ncclRecv with KAMR_HALO=nccl in the environment. */
int  kamr_comm_unique_id(void* id128);
int  kamr_comm_init(kamr_ctx* ctx, const void* id128);

/* re-flatten: after initialize / restart / every amr_recover! (Solver/AMR.jl:54).
 * On a mesh with peers (n_peer > 0) this call is COLLECTIVE over the communicator, as amr_recover! is over MPI: every rank
 * must have called kamr_comm_init first and must call kamr_upload_topology too (the ranks swap their slope-halo
 * schedule and, for the one-sided halo, the CUDA IPC handles and ghost offsets of their arrays); a mesh with peers
 * without a communicator is refused.  All index arrays are range-checked; a malformed mesh is an error, not a crash. */
int  kamr_upload_topology(kamr_ctx* ctx, const kamr_mesh* mesh);
/* state in host layout; df for all n_cell cells (solid-neighbour blocks may be anything), w and prim for local cells
 * [n_local*(DIM+2)].  NULL pointers are skipped.  On a mesh with peers the ghost blocks of df are copied as well and are
 * then replaced by the owners' values with kamr_exchange_df; since the one-sided halo has no rendezvous, a peer that is
 * already past its own upload may have stored its mirrors first — pass the owners' values in the ghost blocks (what
 * GhostPsData holds after data_exchange!), or separate the ranks' upload and exchange calls by a barrier. */
int  kamr_upload_state(kamr_ctx* ctx, const double* df, const double* w, const double* prim);
/* optional: upload sdf / vs flux / macro flux (restart in the middle of a step) */
int  kamr_upload_aux(kamr_ctx* ctx, const double* sdf, const double* flux, const double* mflux);
int  kamr_download_state(kamr_ctx* ctx, uint32_t mask, double* df, double* sdf, double* flux,
                         double* w, double* prim, double* qf, double* sw, double* mflux);

/* Selective transfer of whole cells, packed on the device: what partition migration ships per migrating cell
 * (Parallel/Partition.jl:339-388: w and df; level / midpoint are host-known statics) and what an adapt event reads of
 * the flagged cells (Physical_space/AMR.jl:633-703).  cells[q] are local cell ids; df holds the cells' VsData.df blocks
 * back to back in the given order (n(c)*NDF doubles each, host layout), w the (DIM+2)-vectors.  NULL pointers are
 * skipped.  kamr_unpack_cells is the receiving side (after kamr_upload_topology of the new partition). */
int  kamr_pack_cells(kamr_ctx* ctx, int32_t n, const int32_t* cells, double* df, double* w);
int  kamr_unpack_cells(kamr_ctx* ctx, int32_t n, const int32_t* cells, const double* df, const double* w);

/* Partition migration device to device (ps_partition!, Parallel/Partition.jl:339-445, 517-610): the w, prim
 * and df of the cells that change rank (and of the cells that stay) cross the re-flatten without touching the host.
 * The statics the reference ships beside them (bound_enc, solid_cell_index, vs levels / midpoints) are the host's and
 * arrive through the new kamr_mesh as before.
 *   kamr_migrate_begin, on the OLD topology, collective over the ranks that exchange cells: cells[q] (local id) goes
 *     to rank dest_rank[q] — this rank included, for the cells it keeps.  Per destination the list order is kept.
 *     src_rank[i] (ascending; this rank included if it keeps cells) announces what arrives: src_cells[i] cells with
 *     src_points[i] velocity points in total (the receive_nums / vs_nums the reference exchanges first,
 *     Partition.jl:300-338).  Payloads travel with ncclSend/ncclRecv over NVLink; kept cells are copied on the device.
 *   kamr_upload_topology of the NEW partition (the staging buffer survives it).
 *   kamr_migrate_finish: arrival q — sources in ascending rank, the sender's list order within a source — becomes
 *     local cell recv_cells[q]; its velocity grid must have the size the sender's had.  prim arrives with w (the
 *     reference recomputes get_prim(w), Partition.jl:639).  Follow with kamr_exchange_df on a mesh with peers. */
int  kamr_migrate_begin(kamr_ctx* ctx, int32_t n_send, const int32_t* cells, const int32_t* dest_rank, int32_t n_src,
                        const int32_t* src_rank, const int32_t* src_cells, const int64_t* src_points);
int  kamr_migrate_finish(kamr_ctx* ctx, int32_t n_recv, const int32_t* recv_cells);

/* update_criterion!(ka), Physical_space/AMR.jl:256-341: the physical-space adaptation sensor on the device.  For every
 * local fluid cell and direction the Löhner estimator of Physical_space/Criteria.jl:25-200 over the primitive state
 * of the two sides (mean of the neighbours' conserved state; ds, 0.75 ds or 1.5 ds by the side's level, AMR.jl:5-157;
 * the coarser side shifted to the cell's transverse position with its sw) with the vorticity estimator in row 2, then
 * the one-cell buffer apply_amr_buffer! (AMR.jl:296-341: an unflagged cell with a flagged fluid face neighbour gets
 * lohner .= 2*threshold; the mirrors' decisions travel to the peers as lohner_flag_exchange! does,
 * Parallel/Ghost.jl:939-978).  threshold = ADAPT_COEFFI_PS.  A SolidNeighbor side enters with w = sw = 0 as in the
 * reference (Boundary/Immersed_boundary.jl:300-302), so donor cells see a density jump there; solid cells get zeros.
 * Precondition: kamr_slope since the last change of state (ps_adaptive_mesh_refinement! runs slope! first,
 * AMR.jl:1109-1111) — the call fails otherwise.  Collective over the communicator when the mesh has peers.
 * lohner_out: [n_local][DIM][DIM+2] (PsData.lohner, column-major as in Julia); sensor_out: [n_local] ps_sensor
 * (Criteria.jl:14-23), all the host's ps_refine_flag / ps_coarsen_flag read (:69-101).  NULL pointers are skipped. */
int  kamr_ps_criterion(kamr_ctx* ctx, double threshold, double* lohner_out, double* sensor_out);

/* Velocity-space adaptation inputs on the device: the per-velocity-point decisions of vs_refine! and
 * vs_coarsen! (Velocity_space/AMR.jl:26-115) evaluated where df and sdf live, so that an adapt event downloads one
 * byte per point and decision instead of df + sdf.  The grid surgery itself (refine_grid_stream!,
 * coarsen_grid_stream!, Velocity_space/Rebuild.jl:44,96) stays on the host. */
typedef struct kamr_vs_adapt {
    int32_t mode;           /* ADAPT_VS_MODE: 0 :lohner (default, Solver/Types.jl:99), 1 contribution */
    int32_t maxlevel;       /* AMR_VS_MAXLEVEL */
    int32_t trees[3];       /* config.vs_trees_num */
    int32_t pad_;
    double  vmin[3], vmax[3];   /* config.quadrature = [vmin1, vmax1, vmin2, vmax2, ...] */
    double  coeff_lohner;   /* ADAPT_COEFFI_VS_LOHNER */
    double  coeff_local;    /* ADAPT_COEFFI_VS_LOCAL */
    double  coeff_global;   /* ADAPT_COEFFI_VS_GLOBAL */
    double  vr_density, vr_energy;   /* Velocity_Resolution after the MPI max (contribution mode only) */
} kamr_vs_adapt;
/* vs_resolution(trees, kinfo) before its two MPI.Allreduce(MAX) (Velocity_space/AMR.jl:139-166): out[0] = max over the
 * local fluid cells of maximum(df) * weight, out[1] = the same for the peculiar-energy density; weight is the volume
 * of a finest-level velocity cell.  The host takes the max over the ranks and passes it back in kamr_vs_adapt. */
int  kamr_vs_resolution(kamr_ctx* ctx, const kamr_vs_adapt* par, double out[2]);
/* refine_flag[p] / coarsen_ok[p] for every velocity point of every local cell (host point order, one byte each):
 * the `refine_flags[c]` of vs_refine! (:56-64) and the `coarsen_ok[c]` of vs_coarsen! (:102-110), both evaluated on
 * the grid and state resident on the device.  The criterion distribution is df + max_d |sdf ds_d|
 * (_criterion_cell!, :8-21) with the slopes resident on the device (those of the last kamr_slope, or of the last step
 * under KAMR_OPT_KEEP_SDF — like the reference, which reads whatever VsData.sdf holds); the call fails if no raw
 * slopes are resident.  :lohner mode: vs_lohner_indicator (Velocity_space/Criteria.jl:259-287) over the same-or-coarser
 * face neighbours of vs_face_neighbor (Velocity_space/Neighbor.jl:181-204; the table is built per distinct velocity
 * grid at the first call after a re-flatten), with local_contribution_{refine,coarsen}_flag (Criteria.jl:19-49);
 * contribution mode: contribution_refine_flag / local && global coarsen flags (Criteria.jl:4-6, 55-84).
 * NOTE: the reference evaluates coarsen_ok AFTER refine_grid_stream! has run; for a cell whose grid that pass changed,
 * the host re-evaluates vs_coarsen! on its own (it needs that cell's df only: refinement copies df to the children and
 * zeroes sdf, Rebuild.jl:69,81) — kamr_pack_cells fetches it.  NULL pointers are skipped. */
int  kamr_vs_criterion(kamr_ctx* ctx, const kamr_vs_adapt* par, uint8_t* refine_flag, uint8_t* coarsen_ok);

/* vs_conserved_correction! (Velocity_space/AMR.jl:120-133): conserved_I_porjection!(vs_data, ps_data.w)
 * (Theory/I-projection.jl:144-159) on the listed local cells — the cells whose velocity grid an adaptation pass changed,
 * after the re-flatten and the upload of their regridded df (kamr_unpack_cells).  The h-component of df is projected
 * onto the cell's w by the Newton iteration of solve_I_projection (the same code CIP_Marching runs every step); solid
 * cells in the list are skipped.  In place; on a mesh with peers follow it with kamr_exchange_df. */
int  kamr_project_cells(kamr_ctx* ctx, int32_t n, const int32_t* cells);

/* options.  KAMR_OPT_KEEP_SDF (default 0): kamr_step keeps the limited slopes r*sdf on the device and
 * writes the reference's raw VsData.sdf only for the cells a kernel reads them from; with the option on,
 * every step also writes the raw sdf of every cell so that kamr_download_state(KAMR_DL_SDF) is valid after
 * kamr_step (the host needs sdf before a velocity-space adapt event, Velocity_space/AMR.jl:9-21).
 * kamr_slope always writes both. */
enum { KAMR_OPT_KEEP_SDF = 1 };
int  kamr_set_option(kamr_ctx* ctx, int32_t option, int32_t value);

/* the hot path */
int  kamr_slope(kamr_ctx* ctx);
int  kamr_flux(kamr_ctx* ctx, double dt);
/* res_out: [2*(DIM+2)] = sumRes | sumAvg of residual_check! (Solver/Finalize.jl:5), local to this rank */
int  kamr_iterate(kamr_ctx* ctx, double dt, int32_t want_residual, double* res_out);
/* slope + flux + iterate with the flux kept on chip (DESIGN.md "fused step") */
int  kamr_step(kamr_ctx* ctx, double dt, int32_t want_residual, double* res_out);
/* halo of df after the update (data_exchange!, Parallel/Ghost.jl:841); called by kamr_iterate/kamr_step
 * internally, exported for host-driven sequences (after kamr_upload_state on a mesh with peers) */
int  kamr_exchange_df(kamr_ctx* ctx);
int  kamr_sync(kamr_ctx* ctx);

/* introspection (tests, benchmark accounting) */
typedef struct kamr_stats {
    int64_t n_phase_local;     /* sum of vs_num over local fluid cells (IO/Check.jl:98) */
    int64_t n_points_total;    /* sum of vs_num over all cells */
    int64_t n_relations;       /* distinct (grid,grid) pair maps built */
    int64_t n_slots;           /* cell-face slots */
    int64_t kernel_launches;   /* kernels launched since creation */
    int64_t device_bytes;      /* device memory owned */
    int64_t halo_bytes_per_step;
    int32_t n_levels;          /* slope sweep phases */
    int32_t fused_cells;       /* cells taking the on-chip flux+update path */
} kamr_stats;
int  kamr_get_stats(kamr_ctx* ctx, kamr_stats* out);
/* per-kernel device timing for roofline accounting (bench.py): while enabled, every kernel
 * launch of the hot path is bracketed by CUDA events on the library's stream.  kamr_profile_read
 * synchronises, writes one record per kernel class that ran since the last read and resets. */
typedef struct kamr_kernel_time {
    char    name[32];
    int64_t launches;
    double  total_ms;
} kamr_kernel_time;
int  kamr_profile_enable(kamr_ctx* ctx, int32_t on);
int  kamr_profile_read(kamr_ctx* ctx, kamr_kernel_time* out, int32_t cap, int32_t* n);
/* pair map of grid ga onto grid gb: start[n_a+1]; returns 1 if identity */
int  kamr_get_pair_map(kamr_ctx* ctx, int32_t ga, int32_t gb, int32_t* start, int32_t cap);
/* cell-face slots of a local cell: writes up to cap records {face, sign}; returns count via *n */
int  kamr_get_cell_slots(kamr_ctx* ctx, int32_t cell, int32_t* face, int32_t* sign, int32_t cap, int32_t* n);

/* test hook: y[i] = the device's exp(x[i]) for x[i] <= 0 as the CAIDVM kernels evaluate it (every exponent on that path
 * is -lambda c^2; Theory/Math.jl:96-143 calls Base.exp).  Lets the tests sweep it against libm in ulps. */
int  kamr_debug_exp_nonpos(kamr_ctx* ctx, const double* x, double* y, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* KAMR_H */
