// TEST INFRASTRUCTURE (oracle side): declarations of the GLPK 4.62 entry points the reference's
// ordering ILP calls (SegmentGraph.cpp:3454-4007).  The ordering ILP is OFF the hot path and is
// never reached by the parity harness; the definitions in glpk_stub.cpp abort if called.
#ifndef SHIM_GLPK_H
#define SHIM_GLPK_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct glp_prob glp_prob;
typedef struct { int msg_lev, br_tech, bt_tech; double tol_int, tol_obj; int tm_lim, out_frq, out_dly; void (*cb_func)(void *, void *); void *cb_info; int cb_size, pp_tech; double mip_gap; int mir_cuts, gmi_cuts, cov_cuts, clq_cuts, presolve, binarize, fp_heur, ps_heur, ps_tm_lim, sr_heur, use_sol; const char *save_sol; int alien, flip; double foo_bar[23]; } glp_iocp;
#define GLP_MIN 1
#define GLP_MAX 2
#define GLP_CV 1
#define GLP_IV 2
#define GLP_BV 3
#define GLP_FR 1
#define GLP_LO 2
#define GLP_UP 3
#define GLP_DB 4
#define GLP_FX 5
#define GLP_MSG_OFF 0
#define GLP_MSG_ERR 1
#define GLP_ON 1
#define GLP_OFF 0
#define GLP_UNDEF 1
#define GLP_FEAS 2
#define GLP_INFEAS 3
#define GLP_NOFEAS 4
#define GLP_OPT 5
#define GLP_EBADB 0x01
#define GLP_ESING 0x02
#define GLP_ECOND 0x03
#define GLP_EBOUND 0x04
#define GLP_EFAIL 0x05
#define GLP_EOBJLL 0x06
#define GLP_EOBJUL 0x07
#define GLP_EITLIM 0x08
#define GLP_ETMLIM 0x09
#define GLP_ENOPFS 0x0A
#define GLP_ENODFS 0x0B
#define GLP_EROOT 0x0C
#define GLP_ESTOP 0x0D
#define GLP_EMIPGAP 0x0E
glp_prob *glp_create_prob(void);
void glp_set_prob_name(glp_prob *, const char *);
void glp_set_obj_dir(glp_prob *, int);
int glp_add_rows(glp_prob *, int);
int glp_add_cols(glp_prob *, int);
void glp_set_row_name(glp_prob *, int, const char *);
void glp_set_col_name(glp_prob *, int, const char *);
void glp_set_row_bnds(glp_prob *, int, int, double, double);
void glp_set_col_bnds(glp_prob *, int, int, double, double);
void glp_set_obj_coef(glp_prob *, int, double);
void glp_set_col_kind(glp_prob *, int, int);
void glp_load_matrix(glp_prob *, int, const int[], const int[], const double[]);
void glp_init_iocp(glp_iocp *);
int glp_intopt(glp_prob *, const glp_iocp *);
int glp_mip_status(glp_prob *);
double glp_mip_col_val(glp_prob *, int);
int glp_free_env(void);
#ifdef __cplusplus
}
#endif
#endif
