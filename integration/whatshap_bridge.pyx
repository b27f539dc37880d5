# cython: language_level=3
# distutils: language = c++
"""Reference-side binding: reads a genuine `whatshap.core.ReadSet` / `Pedigree` straight from the C++ objects
behind them into the flat arrays of `whmec_problem` (include/whmec.h) -- the stub a WhatsHap maintainer would add
next to `whatshap/readselect.pyx` (which reaches into `ReadSet.thisptr` the same way, readselect.pyx:244).

It is compiled AGAINST a WhatsHap source tree (its `whatshap/*.pxd` and `src/*.h`; nothing of it is copied) and
resolves the C++ methods at import time from `whatshap.core`, which WhatsHap loads with RTLD_GLOBAL
(whatshap/__init__.py:10-15) -- exactly how `whatshap.readselect` links.  Build: integration/build_bridge.py.
`whatshap_b200.adapters.flatten_objects` uses it when importable and falls back to the public Python API
(one `Variant` object per entry) otherwise.
"""
import numpy as np

from libc.stdint cimport int32_t, uint8_t, uint32_t, uint64_t
from whatshap cimport cpp
from whatshap.core cimport ReadSet


def readset_to_csr(ReadSet readset):
    """(read_off u64[n+1], ent_pos i32, ent_allele u8, ent_quality u32, sample_id i32[n], source_id i32[n]);
    entries in the order the reads store them (Read::getPosition / getAllele / getVariantQuality, src/read.cpp:104-140)."""
    cdef cpp.ReadSet* rs = readset.thisptr
    cdef int n = rs.size()
    cdef int i, j, count
    cdef cpp.Read* read
    cdef uint64_t total = 0
    off = np.zeros(n + 1, np.uint64)
    cdef uint64_t[::1] off_v = off
    for i in range(n):
        total += rs.get(i).getVariantCount()
        off_v[i + 1] = total
    pos = np.empty(total, np.int32)
    allele = np.empty(total, np.uint8)
    quality = np.empty(total, np.uint32)
    sample = np.empty(n, np.int32)
    source = np.empty(n, np.int32)
    cdef int32_t[::1] pos_v = pos
    cdef uint8_t[::1] allele_v = allele
    cdef uint32_t[::1] quality_v = quality
    cdef int32_t[::1] sample_v = sample
    cdef int32_t[::1] source_v = source
    cdef uint64_t e = 0
    for i in range(n):
        read = rs.get(i)
        count = read.getVariantCount()
        sample_v[i] = read.getSampleID()
        source_v[i] = read.getSourceID()
        for j in range(count):
            pos_v[e] = read.getPosition(j)
            allele_v[e] = <uint8_t>read.getAllele(j)
            quality_v[e] = <uint32_t>read.getVariantQuality(j)
            e += 1
    return off, pos, allele, quality, sample, source
