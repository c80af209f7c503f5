// Version / error-string entry points of the C ABI.
#include "host_common.h"
#include "../../include/tricolo_b200.h"

extern "C" int tcl_version(void) { return TCL_ABI_VERSION; }
extern "C" const char* tcl_last_error_string(void) { return tcl::last_error_buf(); }
