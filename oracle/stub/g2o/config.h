/* Stub of the cmake-generated g2o/config.h (reference: config.h.in:1-23), only so that the
 * vendored CSparse + csparse_helper.cpp under /root/reference compile in place into oracle/_ref/.
 * Test infrastructure only. */
#ifndef G2O_CONFIG_H
#define G2O_CONFIG_H
#define G2O_HAVE_CSPARSE 1
#endif
