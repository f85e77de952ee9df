"""ctypes loader for libehyb.so (the C ABI declared in include/*.h).

There is no fallback: if the shared library is missing or a symbol is absent the import
fails loudly.  The library is built in-tree by __graft_entry__.build() / csrc/Makefile.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB_PATH = Path(os.environ.get("EHYB_LIB", PKG / "lib" / "libehyb.so"))

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_i16_p = C.POINTER(C.c_int16)
c_u32_p = C.POINTER(C.c_uint32)
c_i64_p = C.POINTER(C.c_int64)


class MatrixCOO(C.Structure):
    """include/spmv.h matrixCOO (reference spmv.h:17-33)."""
    _fields_ = [
        ("totalNum", C.c_int), ("dimension", C.c_int), ("maxCol", C.c_int), ("nParts", C.c_int),
        ("vectorCacheSize", C.c_uint16), ("kernelPerPart", C.c_int16),
        ("rowIdx", c_int_p), ("numInRow", c_int_p), ("numInRow2", c_int_p),
        ("I", c_int_p), ("J", c_int_p), ("V", c_dbl_p), ("diag", c_dbl_p),
        ("partBoundary", c_int_p), ("reorderList", c_int_p),
    ]


class MatrixEHYB(C.Structure):
    """include/spmv.h matrixEHYB (reference spmv.h:35-63 + appended b200)."""
    _fields_ = [
        ("dimension", C.c_int), ("nParts", C.c_int), ("vectorCacheSize", C.c_int16),
        ("kernelPerPart", C.c_int), ("numOfRowER", C.c_int), ("warpIdxER_d", c_int_p),
        ("reorderList", c_int_p), ("reorderListER", c_int_p),
        ("widthVecBlockELL", c_i16_p), ("biasVecBlockELL", c_int_p),
        ("colBlockELL", c_i16_p), ("valBlockELL", c_dbl_p), ("partBoundary", c_int_p),
        ("widthVecER", c_i16_p), ("rowVecER", c_int_p), ("biasVecER", c_int_p),
        ("colER", c_int_p), ("valER", c_dbl_p), ("outER", c_dbl_p),
        ("nLongVec", C.c_int), ("longVecBoundary", c_int_p), ("longVecRow", c_int_p),
        ("longVecCol", c_int_p), ("longVecVal", c_dbl_p),
        ("b200", C.c_void_p),
    ]


class DeviceInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("sm_count", C.c_int), ("smem_optin_bytes", C.c_int),
        ("smem_per_sm_bytes", C.c_int), ("l2_bytes", C.c_int), ("cc_major", C.c_int),
        ("cc_minor", C.c_int), ("max_persist_l2_bytes", C.c_int), ("hbm_bytes", C.c_size_t),
        ("name", C.c_char * 64),
    ]


class Plan(C.Structure):
    _fields_ = [("nParts", C.c_int), ("W", C.c_int), ("ctasPerPart", C.c_int), ("threads", C.c_int),
                ("ctasPerSM", C.c_int)]


class LayoutOpts(C.Structure):
    _fields_ = [("W", C.c_int), ("ctasPerPart", C.c_int), ("er_fill", C.c_double),
                ("long_row_threshold", C.c_int), ("ncols", C.c_int64), ("halo_in_overflow", C.c_int), ("cache_cap", C.c_int),
                ("min_coverage", C.c_double)]


class SliceDesc(C.Structure):
    _fields_ = [("off256", C.c_uint32), ("w", C.c_uint16), ("wr", C.c_uint16)]


class PartDesc(C.Structure):
    _fields_ = [("rowStart", C.c_int32), ("rowEnd", C.c_int32), ("sliceStart", C.c_int32), ("sliceEnd", C.c_int32),
                ("cacheStart", C.c_int32), ("cacheCount", C.c_int32), ("reserved", C.c_int32 * 2)]


class LayoutView(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("ncols", C.c_int64), ("nnz", C.c_int64),
        ("nParts", C.c_int32), ("W", C.c_int32), ("ctasPerPart", C.c_int32), ("nSlices", C.c_int32),
        ("parts", C.POINTER(PartDesc)), ("slices", C.POINTER(SliceDesc)), ("blob", C.c_void_p),
        ("blobBytes", C.c_int64), ("nOverflow", C.c_int64),
        ("ovfRow", C.POINTER(C.c_int32)), ("ovfCol", C.POINTER(C.c_int32)), ("ovfVal", c_dbl_p),
        ("cacheCols", C.POINTER(C.c_int32)), ("cacheTotal", C.c_int64), ("cacheMax", C.c_int32), ("haloInOverflow", C.c_int32),
        ("nnzEll", C.c_int64), ("nnzRemInSlice", C.c_int64), ("nnzOverflow", C.c_int64),
        ("padEll", C.c_int64), ("padRem", C.c_int64), ("nLongRows", C.c_int64),
        ("algBytes", C.c_int64), ("formatBytes", C.c_int64),
    ]


class PcgOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("rtol", C.c_double), ("check_every", C.c_int)]


class PcgResult(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("rel_residual", C.c_double), ("true_rel_residual", C.c_double),
                ("ms", C.c_float)]


class SessionOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("threads", C.c_int), ("use_graph", C.c_int),
                ("l2_persist_x", C.c_int), ("halo_cols", C.c_int64), ("kernel", C.c_int)]


# every symbol the headers declare; checked at load time (tests/test_abi.py re-checks against
# the header text)
EXPORTS = [
    # spmv.h / kernel.h / convert.h / reordering.h (the reference's entry points)
    "spmvGPuEHYB", "spmvGPuEHYB_layout", "matrixVectorEHYB", "matrixVectorEHYB_small", "COO2EHYB", "EHYBfreeHost",
    "matrixReorder", "matrixReorder_unsym", "vectorReorder", "vectorRecover",
    # mmio.h
    "mm_read_banner", "mm_read_mtx_crd_size", "mm_read_mtx_array_size", "mm_write_banner",
    "mm_write_mtx_crd_size", "mm_write_mtx_array_size", "mm_is_valid", "mm_typecode_to_str",
    "mm_read_mtx_crd_entry", "mm_read_mtx_crd_data", "mm_write_mtx_crd", "mm_read_unsymmetric_sparse",
    # ehyb.h
    "ehyb_last_error", "ehyb_version", "ehyb_set_host_threads", "ehyb_get_host_threads", "ehyb_device_count", "ehyb_device_query", "ehyb_device_info_b200",
    "ehyb_plan", "ehyb_plan_kernel", "ehyb_plan_reference", "ehyb_build_graph", "ehyb_set_partitioner", "ehyb_partition_graph",
    "ehyb_reorder_with_partition", "ehyb_reorder", "ehyb_partition_blocks", "ehyb_free_host",
    "ehyb_layout_build", "ehyb_layout_build_csr", "ehyb_layout_get", "ehyb_layout_to_reference",
    "ehyb_layout_free", "ehyb_layout_save", "ehyb_layout_load", "ehyb_cache_save", "ehyb_cache_load", "ehyb_cache_set_options_tag",
    "ehyb_session_opts_default", "ehyb_upload", "ehyb_spmv", "ehyb_spmv_host",
    "ehyb_spmv_host_batch", "ehyb_session_vectors", "ehyb_plan_auto", "ehyb_set_x", "ehyb_get_y", "ehyb_time_spmv", "ehyb_time_spmv_flushed",
    "ehyb_launches_per_spmv", "ehyb_pcg_opts_default", "ehyb_pcg_solve", "ehyb_session_size", "ehyb_session_kernel", "ehyb_trace_read", "ehyb_sync", "ehyb_stream", "ehyb_free", "ehyb_describe",
    "ehyb_gen_lower", "ehyb_coo_from_lower", "ehyb_coo_from_general", "ehyb_gen_rmat", "ehyb_x_reference",
    "ehyb_read_mtx", "ehyb_write_mtx", "ehyb_coo_free",
    "ehyb_mg_local_build", "ehyb_mg_local_halo", "ehyb_mg_local_set_send", "ehyb_mg_local_graph",
    "ehyb_mg_local_finish", "ehyb_mg_local_view", "ehyb_mg_local_free", "ehyb_mg_unique_id",
    "ehyb_mg_session_create", "ehyb_mg_session_handle", "ehyb_mg_spmv", "ehyb_mg_time_spmv", "ehyb_mg_spmv_host_batch",
    "ehyb_mg_p2p_supported", "ehyb_mg_session_create_p2p", "ehyb_mg_p2p_export", "ehyb_mg_p2p_connect", "ehyb_mg_p2p_connect_local",
    "ehyb_mg_status", "ehyb_mg_launches_per_spmv", "ehyb_mg_allreduce_sum", "ehyb_mg_spmv_dot_supported", "ehyb_mg_spmv_dot",
    "ehyb_mg_pcg_solve", "ehyb_mg_session_ranks", "ehyb_spmv_dot_supported", "ehyb_spmv_dot", "ehyb_spmv_dot_host", "ehyb_session_device",
    "ehyb_mg_session_free", "ehyb_gen_stencil27_rows", "ehyb_host_alloc_pinned", "ehyb_host_free_pinned",
    "ehyb_session_info",
    "ehyb_layout_builder_begin", "ehyb_layout_builder_add", "ehyb_layout_builder_finish", "ehyb_layout_builder_abort",
    "ehyb_grid_brick_graph", "ehyb_partition_graph_weighted", "ehyb_partition_graph_hier", "ehyb_set_partition_pieces", "ehyb_get_partition_pieces", "ehyb_grid_decomp_create", "ehyb_grid_decomp_info",
    "ehyb_grid_decomp_free", "ehyb_mg_grid_build", "ehyb_mg_local_natural_ids", "ehyb_grid_natural_ids", "ehyb_grid_rows",
]

_lib = None


def load(path: Path | None = None, check_exports: bool = True) -> C.CDLL:
    """Load libehyb.so; raises if it is missing or incomplete (no CPU fallback exists)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise ImportError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {PKG / 'csrc'}`; the engine has no fallback path")
    lib = C.CDLL(str(p))
    if check_exports:
        missing = [s for s in EXPORTS if not hasattr(lib, s)]
        if missing:
            raise ImportError(f"{p} does not export: {', '.join(missing)}")
    lib.ehyb_last_error.restype = C.c_char_p
    lib.ehyb_version.restype = C.c_char_p
    if hasattr(lib, "ehyb_session_kernel"):
        lib.ehyb_session_kernel.restype = C.c_char_p
    if hasattr(lib, "ehyb_stream"):
        lib.ehyb_stream.restype = C.c_void_p
    if path is None:
        _lib = lib
    return lib


class EhybError(RuntimeError):
    pass


def check(lib, rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.ehyb_last_error()
        raise EhybError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
