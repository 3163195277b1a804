/* Stand-in for <cblas.h> -- TEST INFRASTRUCTURE.  The reference links "-lblas" (partapp.pro:80) without pinning an
 * implementation; cblas_sdot here is the Netlib reference order (ascending index, fp32 multiply then add), the same
 * convention the oracle states in its header.  Defined in oracle/ref_core.cpp. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
float cblas_sdot(const int n, const float *x, const int incx, const float *y, const int incy);
#ifdef __cplusplus
}
#endif
