# cython: language_level=3
# Drop-in replacement for giggles/readselect.pyx (textually included by giggles/core.pyx:509).
#
# Same public function, signature, return type and error as the reference's readselection()
# (giggles/readselect.pyx:231-271); the selection itself runs in gg_readselection()
# (include/giggles_b200.h, giggles_b200/csrc/readselect.cpp), which reproduces the reference's
# result exactly, ties included.  giggles/priorityqueue.pyx, coverage.py and graph.py are no
# longer needed by this function.
from libc.stdint cimport uint32_t, uint64_t, int32_t
from libcpp.vector cimport vector


cdef extern from "giggles_b200.h":
	ctypedef struct gg_readset_view:
		uint32_t n_reads
		const uint64_t *read_variant_offset
		const int32_t *variant_position
		const int32_t *variant_quality
		const int32_t *read_source_id
	int gg_readselection(const gg_readset_view *rs, uint32_t max_cov, const int32_t *preferred_source_ids,
		uint32_t n_preferred, int bridging, uint32_t *selected, uint32_t *n_selected, char *err, size_t errlen)


def readselection(ReadSet pyreadset, max_cov, preferred_source_ids=None, bridging=True):
	'''Return the selected read indices which do not violate the maximal coverage.'''
	cdef cpp.ReadSet* readset = pyreadset.thisptr
	assert readset != NULL
	cdef vector[uint64_t] off
	cdef vector[int32_t] pos
	cdef vector[int32_t] qual
	cdef vector[int32_t] src
	cdef vector[int32_t] pref
	cdef vector[uint32_t] selected
	cdef cpp.Read* read
	cdef int n = readset.size()
	cdef int k, i
	cdef uint32_t n_selected = 0
	cdef char err[512]
	cdef gg_readset_view view
	off.push_back(0)
	for k in range(n):
		read = readset.get(k)
		for i in range(read.getVariantCount()):
			pos.push_back(read.getPosition(i))
			qual.push_back(read.getQuality(i))
		off.push_back(pos.size())
		src.push_back(read.getSourceID())
	if preferred_source_ids is not None:
		for s in preferred_source_ids:
			pref.push_back(s)
	selected.resize(max(n, 1))
	view.n_reads = n
	view.read_variant_offset = off.data()
	view.variant_position = pos.data()
	view.variant_quality = qual.data()
	view.read_source_id = src.data()
	err[0] = 0
	rc = gg_readselection(&view, max_cov, pref.data() if preferred_source_ids is not None else NULL, pref.size(),
		1 if bridging else 0, selected.data(), &n_selected, err, 512)
	if rc == 2:
		raise ValueError(err.decode())
	if rc != 0:
		raise RuntimeError(err.decode())
	return set(selected[k] for k in range(n_selected))
