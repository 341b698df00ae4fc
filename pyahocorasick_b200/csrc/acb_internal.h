/* internal glue between acb_host.cpp and acb_device.cu (not part of the ABI) */
#ifndef ACB_INTERNAL_H_INCLUDED
#define ACB_INTERNAL_H_INCLUDED

#include "../../include/acb200.h"

#ifdef __cplusplus
extern "C" {
#endif
/* printf-style; stores a thread-local message returned by acb_last_error() */
void acb_set_error(const char *fmt, ...);
#ifdef __cplusplus
}
#endif

#endif
