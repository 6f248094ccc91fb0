/* kamr_oracle.h — CPU oracle (test infrastructure only; see kamr_oracle.c header). */
#ifndef KAMR_ORACLE_H
#define KAMR_ORACLE_H
#include <stdint.h>
#include "../include/kamr.h"
#ifdef __cplusplus
extern "C" {
#endif
/* state in the host layout of include/kamr.h */
typedef struct orc_state {
    double *df, *sdf, *flux;      /* per point */
    double *w, *prim, *mflux;     /* per cell x (DIM+2) */
    double *qf;                   /* per cell x DIM */
    double *sw;                   /* per cell x (DIM+2) x DIM */
} orc_state;

void   orc_get_prim(int D, const double* w, double gamma, double* prim);
void   orc_get_conserved(int D, const double* prim, double gamma, double* w);
double orc_get_tau(int D, const double* prim, double mu, double omega);
void   orc_discrete_maxwell(int D, int K, int n, const double* vmid, const double* prim, double Kin, double* F);
void   orc_shakhov_part(int D, int K, int n, const double* vmid, const double* F, const double* prim,
                        const double* qf, double Pr, double Kin, double* Fp);
void   orc_micro_to_macro(int D, int K, int n, const double* vmid, const double* df, const double* weight, double* w);
void   orc_heat_flux(int D, int K, int n, const double* vmid, const double* df, const double* prim,
                     const double* weight, double* q);
int    orc_pair_map(int D, int n_a, const int8_t* lev_a, int n_b, const int8_t* lev_b, int32_t* start);

int    orc_small_solve(int M, double* A, double* b, double* x);
int    orc_solve_I_projection(int D, int n, const double* vm, double* f, const double* W, const double* wt,
                              double* lam);

int orc_slope_level(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, int Lv, int transverse);
int orc_macro_slope(const kamr_config* cfg, const kamr_mesh* m, orc_state* st);
int orc_slope(const kamr_config* cfg, const kamr_mesh* m, orc_state* st);
int orc_ib_solid_cells(const kamr_config* cfg, const kamr_mesh* m, orc_state* st);
int orc_ib_solid_neighbors(const kamr_config* cfg, const kamr_mesh* m, orc_state* st);
int orc_flux(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, double dt);
int orc_iterate(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, double dt, int want_residual,
                double* res_out);
int orc_step(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, double dt, int want_residual,
             double* res_out);
/* update_criterion!(ka) (Physical_space/AMR.jl:256-341): Löhner sensor of every local cell + the one-cell buffer.
 * st holds w, prim and the macro slopes sw of a finished orc_slope.  ghost_flag NULL on one rank. */
int orc_ps_criterion(const kamr_config* cfg, const kamr_mesh* m, const orc_state* st, double threshold,
                     const int32_t* ghost_flag, double* lohner_out, double* sensor_out, int32_t* flag_out);
/* velocity-space adaptation inputs (Velocity_space/AMR.jl:26-166, Criteria.jl, Neighbor.jl); flags per local point */
int orc_vs_face_neighbors(int D, int n, const double* vmid, const int8_t* level, const kamr_vs_adapt* par, int32_t* out);
int orc_vs_resolution(const kamr_config* cfg, const kamr_mesh* m, const orc_state* st, const kamr_vs_adapt* par,
                      double* out);
int orc_vs_criterion(const kamr_config* cfg, const kamr_mesh* m, const orc_state* st, const kamr_vs_adapt* par,
                     uint8_t* refine_flag, uint8_t* coarsen_ok);
int orc_project_cells(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, int n_list, const int32_t* list);
/* test hook: order of the Newton sums of solve_I_projection over the velocity points (0 forward, 1 reverse) */
void orc_set_cip_sum_order(int reverse);
#ifdef __cplusplus
}
#endif
#endif
