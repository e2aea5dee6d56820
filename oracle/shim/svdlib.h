/* TEST INFRASTRUCTURE — empty stand-in for SVDLIBC's svdlib.h (library absent here).
 * The only reference path that needs SVDLIBC is --mf_method sgdparsvd, which is out of
 * scope; the functions of svdFrmsvdlib.h are provided as aborting link
 * stubs in ref_stubs.cpp. */
#ifndef MFB_ORACLE_SVDLIB_SHIM_H
#define MFB_ORACLE_SVDLIB_SHIM_H
#endif
