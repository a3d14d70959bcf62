// TEST INFRASTRUCTURE (oracle side): abort-on-call GLPK stubs.  The parity harness stops before
// Ordering() (SegmentGraph.cpp:3236), so none of these is ever executed; they exist to link.
#include <cstdio>
#include <cstdlib>
#include "glpk.h"
static void no_glpk(const char *f) { fprintf(stderr, "[oracle/_ref] GLPK stub %s called: the ordering ILP is out of scope for the hot-path harness\n", f); abort(); }
extern "C" {
glp_prob *glp_create_prob(void) { no_glpk("glp_create_prob"); return 0; }
void glp_set_prob_name(glp_prob *, const char *) { no_glpk("glp_set_prob_name"); }
void glp_set_obj_dir(glp_prob *, int) { no_glpk("glp_set_obj_dir"); }
int glp_add_rows(glp_prob *, int) { no_glpk("glp_add_rows"); return 0; }
int glp_add_cols(glp_prob *, int) { no_glpk("glp_add_cols"); return 0; }
void glp_set_row_name(glp_prob *, int, const char *) { no_glpk("glp_set_row_name"); }
void glp_set_col_name(glp_prob *, int, const char *) { no_glpk("glp_set_col_name"); }
void glp_set_row_bnds(glp_prob *, int, int, double, double) { no_glpk("glp_set_row_bnds"); }
void glp_set_col_bnds(glp_prob *, int, int, double, double) { no_glpk("glp_set_col_bnds"); }
void glp_set_obj_coef(glp_prob *, int, double) { no_glpk("glp_set_obj_coef"); }
void glp_set_col_kind(glp_prob *, int, int) { no_glpk("glp_set_col_kind"); }
void glp_load_matrix(glp_prob *, int, const int[], const int[], const double[]) { no_glpk("glp_load_matrix"); }
void glp_init_iocp(glp_iocp *) { no_glpk("glp_init_iocp"); }
int glp_intopt(glp_prob *, const glp_iocp *) { no_glpk("glp_intopt"); return 0; }
int glp_mip_status(glp_prob *) { no_glpk("glp_mip_status"); return 0; }
double glp_mip_col_val(glp_prob *, int) { no_glpk("glp_mip_col_val"); return 0; }
int glp_free_env(void) { no_glpk("glp_free_env"); return 0; }
}
