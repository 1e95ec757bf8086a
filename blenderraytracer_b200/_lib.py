"""ctypes binding of libbrt.so (include/brt.h).  No compute happens in Python and there is no CPU fallback:
if the CUDA library is missing the import of this module raises, and without a GPU ``brt_create`` fails."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRT_LIBBRT") or os.path.join(_HERE, "libbrt.so")   # BRT_LIBBRT: kernel-tuning experiments only

BRT_OK, BRT_E_INVALID, BRT_E_CUDA, BRT_E_PARSE, BRT_E_NOSCENE, BRT_E_CANCELLED, BRT_E_NOMEM, BRT_E_STATE = 0, -1, -2, -3, -4, -5, -6, -7
OBJ_SPHERE, OBJ_PLANE, OBJ_BOX, OBJ_TRIANGLE, OBJ_MESH = range(5)
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_EMISSIVE = range(4)
LIGHT_POINT, LIGHT_DIRECTIONAL = 0, 1
BG = {"gradient": 0, "solid": 1, "hdri": 2, "procedural_sky": 3}
AA = {"none": 0, "supersampling": 1, "stochastic": 2}
AA_CENTER = 3
TONEMAP = {"reinhard": 0, "aces": 1, "linear": 2}
CAM_PERSPECTIVE, CAM_ORTHOGRAPHIC, CAM_OTHER = 0, 1, 2
SAMPLER = {"fast": 0, "reference": 1}
INTEGRATOR = {"auto": 0, "megakernel": 1, "wavefront": 2}
ACCEL = {"auto": 0, "brute": 1, "bvh": 2}

d3 = C.c_double * 3


class brt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("texture", C.c_int32), ("color", d3), ("param", C.c_double)]


TEX = {"solid": 0, "checker": 1, "noise": 2, "marble": 3, "wood": 4}


class brt_texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("odd", d3), ("even", d3), ("scale", C.c_double), ("perm", C.c_uint8 * 256)]


class brt_object(C.Structure):
    _fields_ = [("type", C.c_int32), ("material", C.c_int32), ("a", d3), ("b", d3), ("c", d3),
                ("first_tri", C.c_int64), ("tri_count", C.c_int64)]


class brt_light(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("v", d3), ("color", d3), ("intensity", C.c_double)]


SCENE_CONSTRUCTED = 1   # brt_scene_desc.flags: rows read from constructed objects (stored as they are)


class brt_scene_desc(C.Structure):
    _fields_ = [("objects", C.POINTER(brt_object)), ("n_objects", C.c_int32), ("flags", C.c_int32),
                ("materials", C.POINTER(brt_material)), ("n_materials", C.c_int32), ("_pad1", C.c_int32),
                ("mesh_triangles", C.POINTER(C.c_double)), ("n_mesh_triangles", C.c_int64),
                ("lights", C.POINTER(brt_light)), ("n_lights", C.c_int32), ("_pad2", C.c_int32),
                ("textures", C.POINTER(brt_texture)), ("n_textures", C.c_int32), ("_pad3", C.c_int32)]


class brt_camera(C.Structure):
    _fields_ = [("look_from", d3), ("look_at", d3), ("vup", d3),
                ("vfov", C.c_double), ("aspect", C.c_double), ("aperture", C.c_double), ("focus_dist", C.c_double),
                ("type", C.c_int32), ("use_derived", C.c_int32),
                ("origin", d3), ("lower_left_corner", d3), ("horizontal", d3), ("vertical", d3), ("u", d3), ("v", d3), ("w", d3),
                ("lens_radius", C.c_double)]


class brt_render_params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("aa_mode", C.c_int32), ("tonemap", C.c_int32), ("exposure", C.c_double), ("gamma", C.c_double),
                ("denoise", C.c_int32), ("_pad0", C.c_int32), ("denoise_strength", C.c_double), ("seed", C.c_uint64),
                ("direct_lighting", C.c_int32), ("sampler", C.c_int32), ("integrator", C.c_int32), ("accel", C.c_int32),
                ("spp_batch", C.c_int32), ("count_tests", C.c_int32), ("refill_threshold", C.c_int32), ("paths_in_flight", C.c_int32), ("preview", C.c_int32), ("bvh_width", C.c_int32)]


class brt_scene_info(C.Structure):
    _fields_ = [("n_objects", C.c_int32), ("n_materials", C.c_int32), ("n_lights", C.c_int32),
                ("n_spheres", C.c_int32), ("n_planes", C.c_int32), ("n_boxes", C.c_int32),
                ("n_triangles", C.c_int64), ("n_bvh_nodes", C.c_int64), ("bvh_depth", C.c_int32), ("bvh_width", C.c_int32),
                ("bvh_build_ms", C.c_double), ("upload_ms", C.c_double), ("upload_bytes", C.c_int64),
                ("bvh_wide_depth", C.c_int32), ("_pad", C.c_int32), ("bvh_wide_build_ms", C.c_double)]


class brt_stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64),
                ("tests_sphere", C.c_uint64), ("tests_plane", C.c_uint64), ("tests_box", C.c_uint64),
                ("tests_tri_a", C.c_uint64), ("tests_tri_b", C.c_uint64), ("tests_tri_c", C.c_uint64), ("tests_aabb", C.c_uint64),
                ("kernel_ms", C.c_double), ("post_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_uint64),
                ("trav_warp_iters", C.c_uint64), ("trav_lane_iters", C.c_uint64), ("trav_alive_lanes", C.c_uint64),
                ("trav_node_issues", C.c_uint64), ("trav_leaf_issues", C.c_uint64), ("trav_leaf_lanes", C.c_uint64),
                ("path_warp_iters", C.c_uint64), ("path_lane_iters", C.c_uint64), ("node_visits", C.c_uint64)]


PROGRESS_CB = C.CFUNCTYPE(None, C.c_double, C.c_void_p)

# name -> (restype, argtypes); every symbol include/brt.h declares
SIGNATURES = {
    "brt_abi_version": (C.c_int, []),
    "brt_version": (C.c_char_p, []),
    "brt_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "brt_destroy": (None, [C.c_void_p]),
    "brt_last_error": (C.c_char_p, [C.c_void_p]),
    "brt_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "brt_scene_load_json": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "brt_scene_load_binary": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "brt_scene_set_flat": (C.c_int, [C.c_void_p, C.POINTER(brt_scene_desc)]),
    "brt_scene_get_flat": (C.c_int, [C.c_void_p, C.POINTER(brt_scene_desc)]),
    "brt_scene_info_get": (C.c_int, [C.c_void_p, C.POINTER(brt_scene_info)]),
    "brt_set_camera": (C.c_int, [C.c_void_p, C.POINTER(brt_camera)]),
    "brt_get_camera": (C.c_int, [C.c_void_p, C.POINTER(brt_camera)]),
    "brt_set_background": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_uint8)]),
    "brt_get_background": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "brt_set_render_params": (C.c_int, [C.c_void_p, C.POINTER(brt_render_params)]),
    "brt_get_render_params": (C.c_int, [C.c_void_p, C.POINTER(brt_render_params)]),
    "brt_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, PROGRESS_CB, C.c_void_p]),
    "brt_cancel": (None, [C.c_void_p]),
    "brt_get_stats": (C.c_int, [C.c_void_p, C.POINTER(brt_stats)]),
    "brt_render_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "brt_resolve_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "brt_reduce_resolve_peers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "brt_stream_synchronize": (C.c_int, [C.c_void_p]),
    "brt_shared_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "brt_shared_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "brt_shared_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "brt_shared_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "brt_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "brt_device_memset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "brt_primary_aov_f32": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "brt_primary_aov_f64": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "brt_eval_background": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_float)]),
    "brt_debug_rng_stream": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_float)]),
    "brt_postprocess_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "brt_eval_texture": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_float)]),
    "brt_measure_fp32_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "brt_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]),
    "brt_device_count": (C.c_int, [C.c_void_p]),
    "brt_peer_alloc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "brt_peer_connect": (C.c_int, [C.c_void_p, C.c_char_p]),
    "brt_peer_render": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "brt_peer_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "brt_peer_image_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "brt_peer_free": (C.c_int, [C.c_void_p]),
}

_lib = None


class BrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbrt error {code}: {msg}")
        self.code = code


def load():
    """dlopen libbrt.so and bind every entry point.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C blenderraytracer_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(ctx, rc):
    if rc != BRT_OK:
        msg = load().brt_last_error(ctx)
        raise BrtError(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc
