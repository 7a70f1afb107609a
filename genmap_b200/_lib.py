"""ctypes binding of libgenmap_b200.so (include/genmap_b200.h).  Fails loudly when the library is
missing: there is no Python or CPU fallback for the compute path."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMB_LIB_PATH") or os.path.join(HERE, "lib", "libgenmap_b200.so")  # env: tuning builds only

GMB_OK, GMB_ERR_ARG, GMB_ERR_UNSUPPORTED, GMB_ERR_CUDA, GMB_ERR_IO, GMB_ERR_NOMEM = 0, -1, -2, -3, -4, -5
GMB_BUILD_WITH_SA, GMB_BUILD_ON_GPU = 1, 2

EXPORTS = ["gmb_last_error", "gmb_version", "gmb_device_count", "gmb_index_build", "gmb_blob_free",
           "gmb_index_build_device", "gmb_blob_save", "gmb_index_open", "gmb_index_from_blob",
           "gmb_index_adopt_device", "gmb_index_close", "gmb_index_get_info", "gmb_map_frequencies",
           "gmb_map_frequencies_range", "gmb_map_frequencies_device", "gmb_index_export_bwt", "gmb_index_export_sa", "gmb_index_set_jump_depth",
           "gmb_index_import_reference", "gmb_map_locations", "gmb_locations_free", "gmb_map_runs", "gmb_runs_free", "gmb_index_replicate",
           "gmb_index_set_plan_text_size", "gmb_progress", "gmb_blob_export_reference"]


class GmbParams(ctypes.Structure):
    _fields_ = [("K", ctypes.c_uint32), ("E", ctypes.c_uint32), ("revcompl", ctypes.c_uint32),
                ("exclude_pseudo", ctypes.c_uint32), ("value_bits", ctypes.c_uint32),
                ("count_fetches", ctypes.c_uint32), ("block_kmers", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class GmbIndexInfo(ctypes.Structure):
    _fields_ = [("n_text", ctypes.c_uint64), ("n_bwt", ctypes.c_uint64), ("n_seq", ctypes.c_uint32),
                ("has_sa", ctypes.c_uint32), ("blob_bytes", ctypes.c_uint64), ("rank_block_bytes", ctypes.c_uint64),
                ("device_blob", ctypes.c_void_p), ("device", ctypes.c_int32), ("alphabet_size", ctypes.c_int32),
                ("jump_table_bytes", ctypes.c_uint64)]


class GmbMapStats(ctypes.Structure):
    _fields_ = [("kernel_ms", ctypes.c_double), ("positions", ctypes.c_uint64),
                ("rank_block_fetches", ctypes.c_uint64), ("jump_table_reads", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint32), ("jump_depth", ctypes.c_uint32),
                ("fetches_by_size", ctypes.c_uint64 * 8), ("thin_paths", ctypes.c_uint64), ("iterations", ctypes.c_uint64),
                ("located_entries", ctypes.c_uint64), ("text_reads", ctypes.c_uint64),
                ("block_kmers", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class GmbLocations(ctypes.Structure):
    _fields_ = [("pos_begin", ctypes.c_uint64), ("pos_end", ctypes.c_uint64), ("n_locations", ctypes.c_uint64),
                ("offsets", ctypes.POINTER(ctypes.c_uint64)), ("loc", ctypes.c_void_p), ("kernel_ms", ctypes.c_double)]


class GmbRuns(ctypes.Structure):
    _fields_ = [("pos_begin", ctypes.c_uint64), ("pos_end", ctypes.c_uint64), ("n_runs", ctypes.c_uint64),
                ("start", ctypes.POINTER(ctypes.c_uint64)), ("value", ctypes.POINTER(ctypes.c_uint16)),
                ("kernel_ms", ctypes.c_double), ("rle_ms", ctypes.c_double)]


class GenmapError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("genmap_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m genmap_b200._build` (needs nvcc); "
                          "genmap_b200 has no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, u64, u32, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    pp = ctypes.POINTER(vp)
    L.gmb_last_error.restype = ctypes.c_char_p
    L.gmb_version.restype = ctypes.c_char_p
    L.gmb_device_count.restype = ci
    L.gmb_index_build.restype = ci
    L.gmb_index_build.argtypes = [vp, vp, u32, u32, ci, pp, ctypes.POINTER(u64)]
    L.gmb_blob_free.argtypes = [vp]
    L.gmb_index_build_device.restype = ci
    L.gmb_index_build_device.argtypes = [vp, vp, u32, u32, ci, pp, ctypes.POINTER(ctypes.c_double)]
    L.gmb_blob_save.restype = ci
    L.gmb_blob_save.argtypes = [vp, u64, ctypes.c_char_p]
    L.gmb_index_open.restype = ci
    L.gmb_index_open.argtypes = [ctypes.c_char_p, ci, pp]
    L.gmb_index_from_blob.restype = ci
    L.gmb_index_from_blob.argtypes = [vp, u64, ci, pp]
    L.gmb_index_adopt_device.restype = ci
    L.gmb_index_adopt_device.argtypes = [vp, u64, ci, pp]
    L.gmb_index_replicate.restype = ci
    L.gmb_index_replicate.argtypes = [vp, ci, pp]
    L.gmb_index_close.restype = ci
    L.gmb_index_close.argtypes = [vp]
    L.gmb_index_get_info.restype = ci
    L.gmb_index_get_info.argtypes = [vp, ctypes.POINTER(GmbIndexInfo)]
    L.gmb_map_frequencies.restype = ci
    L.gmb_map_frequencies.argtypes = [vp, ctypes.POINTER(GmbParams), u64, u64, vp, u32, vp, u64, vp, u32, vp,
                                      ctypes.POINTER(GmbMapStats)]
    L.gmb_map_frequencies_range.restype = ci
    L.gmb_map_frequencies_range.argtypes = [vp, ctypes.POINTER(GmbParams), u64, u64, vp, u32, vp, u64, vp, u32,
                                            u64, u64, vp, ctypes.POINTER(GmbMapStats)]
    L.gmb_index_set_jump_depth.restype = ci
    L.gmb_index_set_jump_depth.argtypes = [vp, ci]
    L.gmb_index_set_plan_text_size.restype = ci
    L.gmb_index_set_plan_text_size.argtypes = [vp, u64]
    L.gmb_progress.restype = ci
    L.gmb_progress.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.gmb_blob_export_reference.restype = ci
    L.gmb_blob_export_reference.argtypes = [vp, u64, ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), u32, ci, u32]
    L.gmb_index_import_reference.restype = ci
    L.gmb_index_import_reference.argtypes = [ctypes.c_char_p, pp, ctypes.POINTER(u64)]
    L.gmb_index_export_sa.restype = ci
    L.gmb_index_export_sa.argtypes = [vp, vp]
    L.gmb_index_export_bwt.restype = ci
    L.gmb_index_export_bwt.argtypes = [vp, ci, vp]
    L.gmb_map_frequencies_device.restype = ci
    L.gmb_map_frequencies_device.argtypes = [vp, ctypes.POINTER(GmbParams), u64, u64, vp, u32, vp, u64, vp, u32,
                                             u64, u64, vp, vp, ctypes.POINTER(GmbMapStats)]
    L.gmb_map_locations.restype = ci
    L.gmb_map_locations.argtypes = [vp, ctypes.POINTER(GmbParams), u64, u64, vp, u32, vp, u64, u64, u64, u64,
                                    ctypes.POINTER(GmbLocations)]
    L.gmb_locations_free.argtypes = [ctypes.POINTER(GmbLocations)]
    L.gmb_map_runs.restype = ci
    L.gmb_map_runs.argtypes = [vp, ctypes.POINTER(GmbParams), u64, u64, vp, u32, vp, u64, vp, u32, u64, u64,
                               ctypes.POINTER(GmbRuns), ctypes.POINTER(GmbMapStats)]
    L.gmb_runs_free.argtypes = [ctypes.POINTER(GmbRuns)]
    _lib = L
    return L


def check(rc):
    if rc != GMB_OK:
        raise GenmapError(rc, lib().gmb_last_error().decode(errors="replace"))
