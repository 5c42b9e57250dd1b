/*
 * oracle/ref_ddsbase_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C-linkage wrappers around the REFERENCE's own PVM/DDS codec so that tests can call it
 * through ctypes.  This file contains no reference code: it includes the reference header
 * from /root/reference/include at build time (-I) and is linked against
 * /root/reference/src/ddsbase.cpp compiled where it lies (see oracle/Makefile, target
 * _ref/libddsbase_ref.so).  Outputs go to oracle/_ref/ only (git-ignored).
 */
#include "ddsbase.h"

extern "C" {

unsigned char* ref_readPVMvolume(const char* fn, unsigned int* w, unsigned int* h, unsigned int* d,
                                 unsigned int* comps, float* sx, float* sy, float* sz)
{
    return readPVMvolume(fn, w, h, d, comps, sx, sy, sz);
}

void ref_writePVMvolume(const char* fn, unsigned char* vol, unsigned int w, unsigned int h,
                        unsigned int d, unsigned int comps, float sx, float sy, float sz,
                        const char* description, const char* courtesy,
                        const char* parameter, const char* comment)
{
    writePVMvolume(fn, vol, w, h, d, comps, sx, sy, sz,
                   (unsigned char*)description, (unsigned char*)courtesy,
                   (unsigned char*)parameter, (unsigned char*)comment);
}

unsigned int ref_checksum(unsigned char* data, unsigned int bytes) { return checksum(data, bytes); }

void ref_free(void* p) { free(p); }

}
