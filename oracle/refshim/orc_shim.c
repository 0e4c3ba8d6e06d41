/*
 * Interpreter for the five Orc opcodes that the reference builds at run time
 * in schromotion8.c:15-167 (convubw, mullw, addw, shrsw, avgub).
 * TEST INFRASTRUCTURE (oracle/_ref build only).  Semantics follow Orc's
 * opcode definitions: word ops wrap at 16 bits, shrsw is arithmetic,
 * avgub(a,b) = (a+b+1)>>1 on unsigned bytes, convubw zero-extends.
 */
#include <stdlib.h>
#include <orc/orc.h>

enum { OP_CONVUBW, OP_MULLW, OP_ADDW, OP_SHRSW, OP_AVGUB };

static int is_param (int v) { return v >= ORC_VAR_P1 && v <= ORC_VAR_P8; }
static int is_const (int v) { return v >= ORC_VAR_C1 && v <= ORC_VAR_C8; }
static int is_temp (int v) { return v >= ORC_VAR_T1 && v <= ORC_VAR_T8; }

static void
shim_exec (OrcExecutor *ex)
{
  OrcProgram *p = ex->program;
  int m = p->is_2d ? ORC_EXECUTOR_M (ex) : 1;
  int n = ex->n;
  int x, y, k;

  for (y = 0; y < m; y++) {
    for (x = 0; x < n; x++) {
      int temps[ORC_N_VARIABLES_SHIM];
      for (k = 0; k < p->n_insns; k++) {
        int vals[2];
        int srcs[2];
        int a;
        int r = 0;
        int d = p->insns[k].d;
        srcs[0] = p->insns[k].s1;
        srcs[1] = p->insns[k].s2;
        for (a = 0; a < 2; a++) {
          int v = srcs[a];
          if (is_param (v)) {
            vals[a] = ex->params[v];
          } else if (is_const (v)) {
            vals[a] = p->const_val[v];
          } else if (is_temp (v)) {
            vals[a] = temps[v];
          } else {
            char *base = (char *) ex->arrays[v] + (size_t) y * ex->params[v];
            if (p->var_size[v] == 1)
              vals[a] = ((uint8_t *) base)[x];
            else
              vals[a] = ((int16_t *) base)[x];
          }
        }
        switch (p->insns[k].op) {
          case OP_CONVUBW: r = (uint8_t) vals[0]; break;
          case OP_MULLW: r = (int16_t) ((int16_t) vals[0] * (int16_t) vals[1]); break;
          case OP_ADDW: r = (int16_t) ((int16_t) vals[0] + (int16_t) vals[1]); break;
          case OP_SHRSW: r = (int16_t) (((int16_t) vals[0]) >> vals[1]); break;
          case OP_AVGUB: r = (uint8_t) (((uint8_t) vals[0] + (uint8_t) vals[1] + 1) >> 1); break;
        }
        if (is_temp (d)) {
          temps[d] = r;
        } else {
          char *base = (char *) ex->arrays[d] + (size_t) y * ex->params[d];
          if (p->var_size[d] == 1)
            ((uint8_t *) base)[x] = (uint8_t) r;
          else
            ((int16_t *) base)[x] = (int16_t) r;
        }
      }
    }
  }
}

OrcProgram *orc_program_new (void) { return calloc (1, sizeof (OrcProgram)); }
void orc_program_set_constant_n (OrcProgram *p, int n) { p->constant_n = n; }
void orc_program_set_2d (OrcProgram *p) { p->is_2d = 1; }
void orc_program_set_name (OrcProgram *p, const char *name) { (void) p; (void) name; }
int orc_program_add_destination (OrcProgram *p, int size, const char *name)
{ int v = ORC_VAR_D1 + p->n_dest++; p->var_size[v] = size; (void) name; return v; }
int orc_program_add_source (OrcProgram *p, int size, const char *name)
{ int v = ORC_VAR_S1 + p->n_src++; p->var_size[v] = size; (void) name; return v; }
int orc_program_add_temporary (OrcProgram *p, int size, const char *name)
{ int v = ORC_VAR_T1 + p->n_temp++; p->var_size[v] = size; (void) name; return v; }
int orc_program_add_parameter (OrcProgram *p, int size, const char *name)
{ int v = ORC_VAR_P1 + p->n_param++; p->var_size[v] = size; (void) name; return v; }
int orc_program_add_constant (OrcProgram *p, int size, int value, const char *name)
{ int v = ORC_VAR_C1 + p->n_const++; p->var_size[v] = size; p->const_val[v] = value; (void) name; return v; }

void
orc_program_append (OrcProgram *p, const char *opcode, int d, int s1, int s2)
{
  int op = -1;
  if (!strcmp (opcode, "convubw")) op = OP_CONVUBW;
  else if (!strcmp (opcode, "mullw")) op = OP_MULLW;
  else if (!strcmp (opcode, "addw")) op = OP_ADDW;
  else if (!strcmp (opcode, "shrsw")) op = OP_SHRSW;
  else if (!strcmp (opcode, "avgub")) op = OP_AVGUB;
  else abort ();
  p->insns[p->n_insns].op = op;
  p->insns[p->n_insns].d = d;
  p->insns[p->n_insns].s1 = s1;
  p->insns[p->n_insns].s2 = s2;
  p->n_insns++;
}

OrcCompileResult
orc_program_compile (OrcProgram *p)
{
  p->code_exec = shim_exec;
  return 0;
}
