/* hd_error.c -- process-global sticky error bitfield and message queue.
 * Same contract as the reference's src/internal/error.c (bits: include/internal/error.h:16-47):
 * API calls OR bits into the state and return it; messages queue up until described. */
#include "hd_internal.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#define HD_MAX_MSGS 64
static uint32_t g_code = 0;
static char    *g_msgs[HD_MAX_MSGS];
static int      g_nmsgs = 0;

void     hd_err_set(uint32_t bits) { g_code |= bits; }
uint32_t hd_err_get(void) { return g_code; }
void     hd_err_reset(void) { g_code = 0; hd_err_clear_msgs(); }
int      hd_err_msg_count(void) { return g_nmsgs; }
const char *hd_err_msg_get(int i) { return (i >= 0 && i < g_nmsgs) ? g_msgs[i] : NULL; }

void hd_err_msg(const char *fmt, ...)
{
   char    buf[2048];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof(buf), fmt, ap);
   va_end(ap);
   for (int i = 0; i < g_nmsgs; i++)
      if (!strcmp(g_msgs[i], buf)) return; /* unique */
   if (g_nmsgs < HD_MAX_MSGS) g_msgs[g_nmsgs++] = strdup(buf);
}

void hd_err_clear_msgs(void)
{
   for (int i = 0; i < g_nmsgs; i++) free(g_msgs[i]);
   g_nmsgs = 0;
}

void hd_err_print_msgs(void)
{
   for (int i = 0; i < g_nmsgs; i++) fprintf(stderr, "--> %s\n", g_msgs[i]);
}

static const struct { uint32_t bit; const char *text; } g_desc[] = {
   {HYPREDRV_ERROR_YAML_INVALID_INDENT, "invalid indentation in YAML input"},
   {HYPREDRV_ERROR_YAML_INVALID_BASE_INDENT, "invalid base indentation in YAML input"},
   {HYPREDRV_ERROR_YAML_INCONSISTENT_INDENT, "inconsistent indentation in YAML input"},
   {HYPREDRV_ERROR_YAML_INVALID_DIVISOR, "missing ':' divisor in YAML input"},
   {HYPREDRV_ERROR_YAML_TREE_NULL, "YAML tree was not built"},
   {HYPREDRV_ERROR_YAML_TREE_INVALID, "YAML tree is invalid"},
   {HYPREDRV_ERROR_YAML_MIXED_INDENT, "tabs and spaces mixed in YAML indentation"},
   {HYPREDRV_ERROR_YAML_INVALID_INDENT_JUMP, "indentation jumps more than one level in YAML input"},
   {HYPREDRV_ERROR_INVALID_KEY, "invalid key(s) in input"},
   {HYPREDRV_ERROR_INVALID_VAL, "invalid value(s) in input"},
   {HYPREDRV_ERROR_UNEXPECTED_VAL, "unexpected value(s) in input"},
   {HYPREDRV_ERROR_MAYBE_INVALID_VAL, "possibly invalid value(s) in input"},
   {HYPREDRV_ERROR_MISSING_KEY, "missing key(s) in input"},
   {HYPREDRV_ERROR_EXTRA_KEY, "extra (unused) key(s) in input"},
   {HYPREDRV_ERROR_MISSING_SOLVER, "missing solver key"},
   {HYPREDRV_ERROR_MISSING_PRECON, "missing preconditioner key"},
   {HYPREDRV_ERROR_MISSING_DOFMAP, "missing dofmap"},
   {HYPREDRV_ERROR_INVALID_SOLVER, "invalid solver"},
   {HYPREDRV_ERROR_INVALID_PRECON, "invalid preconditioner"},
   {HYPREDRV_ERROR_FILE_NOT_FOUND, "file not found"},
   {HYPREDRV_ERROR_FILE_UNEXPECTED_ENTRY, "unexpected entry in file"},
   {HYPREDRV_ERROR_UNKNOWN_HYPREDRV_OBJ, "unknown HYPREDRV object"},
   {HYPREDRV_ERROR_HYPREDRV_NOT_INITIALIZED, "HYPREDRV is not initialized"},
   {HYPREDRV_ERROR_UNKNOWN_TIMING, "unknown timing region"},
   {HYPREDRV_ERROR_HYPRE_INTERNAL, "device solver backend (hdk) reported an error"},
   {HYPREDRV_ERROR_MISSING_LIB, "missing library"},
   {HYPREDRV_ERROR_ALLOCATION, "allocation failure"},
   {HYPREDRV_ERROR_OUT_OF_BOUNDS, "out of bounds access"},
   {HYPREDRV_ERROR_UNKNOWN, "unknown error"},
};

void hd_err_describe(uint32_t code)
{
   if (!code) return;
   fprintf(stderr, "HYPREDRIVE Failure!!!\n");
   for (size_t i = 0; i < sizeof(g_desc) / sizeof(g_desc[0]); i++)
      if (code & g_desc[i].bit) fprintf(stderr, "--> error 0x%08x: %s\n", g_desc[i].bit, g_desc[i].text);
   hd_err_print_msgs();
   fflush(stderr);
}

/* ---- the reference's internal error-module entry points that its own unit tests bind directly
 * (include/internal/error.h:56-79 of the reference; e.g. tests/test_setmatrix_from_csr.c:255):
 * exported so those tests link against this library unmodified ---- */
void     hypredrv_ErrorCodeSet(uint32_t bits) { hd_err_set(bits); }
uint32_t hypredrv_ErrorCodeGet(void) { return hd_err_get(); }
int      hypredrv_ErrorCodeActive(void) { return hd_err_get() != 0; }
void     hypredrv_ErrorCodeReset(uint32_t bits) { g_code &= ~bits; }
void     hypredrv_ErrorCodeResetAll(void) { g_code = 0; }
void     hypredrv_ErrorStateReset(void) { hd_err_reset(); }
void     hypredrv_ErrorMsgPrint(void) { hd_err_print_msgs(); }
void     hypredrv_ErrorMsgClear(void) { hd_err_clear_msgs(); }
