/*
 * Minimal stand-in for liborc's <orc/orc.h>, used ONLY to compile the
 * unmodified reference sources (under /root/reference) into oracle/_ref/
 * with -DDISABLE_ORC.  TEST INFRASTRUCTURE: never linked into the product.
 *
 * liborc (orc-0.4 >= 0.4.16, configure.ac:51-56 of the reference) is not
 * installable offline.  With DISABLE_ORC every orc_* kernel of the reference
 * is plain C (schroorc-dist.c), so all that is needed from liborc is
 *   - the integer typedefs / unions,
 *   - orc_memcpy / orc_memset / orc_init,
 *   - the *runtime program* API that schromotion8.c:15-299 uses to build five
 *     tiny kernels (opcodes convubw mullw addw shrsw avgub).  We provide an
 *     interpreter for exactly those opcodes (orc_shim.c), with Orc's published
 *     opcode semantics (16-bit wrap for addw/mullw, arithmetic shrsw,
 *     avgub = (a+b+1)>>1).
 */
#ifndef ORACLE_ORC_SHIM_H
#define ORACLE_ORC_SHIM_H

#include <stdint.h>
#include <string.h>

#ifndef _ORC_INTEGER_TYPEDEFS_
#define _ORC_INTEGER_TYPEDEFS_
typedef int8_t orc_int8;
typedef int16_t orc_int16;
typedef int32_t orc_int32;
typedef int64_t orc_int64;
typedef uint8_t orc_uint8;
typedef uint16_t orc_uint16;
typedef uint32_t orc_uint32;
typedef uint64_t orc_uint64;
#define ORC_UINT64_C(x) UINT64_C(x)
typedef union { orc_int16 i; orc_int8 x2[2]; } orc_union16;
typedef union { orc_int32 i; float f; orc_int16 x2[2]; orc_int8 x4[4]; } orc_union32;
typedef union { orc_int64 i; double f; orc_int32 x2[2]; float x2f[2]; orc_int16 x4[4]; } orc_union64;
#endif
#ifndef ORC_RESTRICT
#define ORC_RESTRICT restrict
#endif

#define orc_memcpy(d, s, n) memmove ((d), (s), (n))
#define orc_memset(d, v, n) memset ((d), (v), (n))
static inline void orc_init (void) { }

/* ---- runtime program API subset ---- */
enum {
  ORC_VAR_D1 = 0, ORC_VAR_D2, ORC_VAR_D3, ORC_VAR_D4,
  ORC_VAR_S1, ORC_VAR_S2, ORC_VAR_S3, ORC_VAR_S4,
  ORC_VAR_S5, ORC_VAR_S6, ORC_VAR_S7, ORC_VAR_S8,
  ORC_VAR_A1, ORC_VAR_A2, ORC_VAR_A3, ORC_VAR_A4,
  ORC_VAR_C1, ORC_VAR_C2, ORC_VAR_C3, ORC_VAR_C4,
  ORC_VAR_C5, ORC_VAR_C6, ORC_VAR_C7, ORC_VAR_C8,
  ORC_VAR_P1, ORC_VAR_P2, ORC_VAR_P3, ORC_VAR_P4,
  ORC_VAR_P5, ORC_VAR_P6, ORC_VAR_P7, ORC_VAR_P8,
  ORC_VAR_T1, ORC_VAR_T2, ORC_VAR_T3, ORC_VAR_T4,
  ORC_VAR_T5, ORC_VAR_T6, ORC_VAR_T7, ORC_VAR_T8,
  ORC_N_VARIABLES_SHIM = 64
};

typedef struct _OrcExecutor OrcExecutor;
typedef struct _OrcProgram OrcProgram;
typedef int OrcCompileResult;
#define ORC_COMPILE_RESULT_IS_SUCCESSFUL(x) ((x) == 0)

struct _OrcProgram {
  int n_insns;
  struct { int op, d, s1, s2; } insns[16];
  int var_size[ORC_N_VARIABLES_SHIM];
  int const_val[ORC_N_VARIABLES_SHIM];
  int constant_n;
  int is_2d;
  int n_const, n_param, n_temp, n_src, n_dest;
  void (*code_exec) (OrcExecutor *);
};

struct _OrcExecutor {
  OrcProgram *program;
  int n;
  int counter1, counter2, counter3;
  void *arrays[ORC_N_VARIABLES_SHIM];
  int params[ORC_N_VARIABLES_SHIM];
  int accumulators[4];
};
/* real liborc keeps m in params[ORC_VAR_A1] */
#define ORC_EXECUTOR_M(ex) ((ex)->params[ORC_VAR_A1])

OrcProgram *orc_program_new (void);
void orc_program_set_constant_n (OrcProgram *p, int n);
void orc_program_set_2d (OrcProgram *p);
void orc_program_set_name (OrcProgram *p, const char *name);
int orc_program_add_destination (OrcProgram *p, int size, const char *name);
int orc_program_add_source (OrcProgram *p, int size, const char *name);
int orc_program_add_temporary (OrcProgram *p, int size, const char *name);
int orc_program_add_parameter (OrcProgram *p, int size, const char *name);
int orc_program_add_constant (OrcProgram *p, int size, int value, const char *name);
void orc_program_append (OrcProgram *p, const char *opcode, int d, int s1, int s2);
OrcCompileResult orc_program_compile (OrcProgram *p);

#endif
