"""ctypes mirror of include/b2rt.h (structs only).  Shared by the product binding (b2rt/__init__.py)
and by the oracle's test wrapper (oracle/orc.py) because both speak the same flat scene ABI."""
import ctypes as C


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("albedo", C.c_float * 3), ("transmittance", C.c_float * 3),
                ("emission", C.c_float * 3), ("ior", C.c_float), ("roughness", C.c_float)]


class Light(C.Structure):
    _fields_ = [("kind", C.c_int32), ("radiance", C.c_float * 3), ("position", C.c_float * 3),
                ("direction", C.c_float * 3), ("dim_x", C.c_float * 3), ("dim_y", C.c_float * 3)]


class SceneDesc(C.Structure):
    _fields_ = [("n_tris", C.c_uint32), ("tri_verts", C.POINTER(C.c_float)), ("tri_normals", C.POINTER(C.c_float)),
                ("tri_material", C.POINTER(C.c_uint32)), ("n_spheres", C.c_uint32),
                ("spheres", C.POINTER(C.c_float)), ("sphere_material", C.POINTER(C.c_uint32)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)), ("n_lights", C.c_uint32),
                ("lights", C.POINTER(Light))]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("c2w", C.c_float * 9), ("hfov_deg", C.c_float), ("vfov_deg", C.c_float)]


class SceneFile(C.Structure):
    """b2rt_scene_file: what b2rt_scene_load / b2rt_load_dae return (storage owned by the library)."""
    _fields_ = [("desc", SceneDesc), ("camera", Camera), ("cam_dir", C.c_float * 3), ("cam_hfov_deg", C.c_float),
                ("cam_vfov_deg", C.c_float), ("bbox", C.c_float * 6), ("storage", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("ns_aa", C.c_uint32), ("max_ray_depth", C.c_uint32), ("ns_area_light", C.c_uint32),
                ("seed", C.c_uint64), ("ray_eps", C.c_float), ("bvh_width", C.c_uint32),
                ("max_leaf_size", C.c_uint32), ("treelet_bytes", C.c_uint32), ("max_wave_paths", C.c_uint32),
                ("median_threshold", C.c_uint32), ("device", C.c_int32), ("sample_first", C.c_uint32),
                ("sample_stride", C.c_uint32), ("bvh_builder", C.c_uint32), ("filter_kind", C.c_uint32),
                ("filter_sigma_r", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("rays_camera", C.c_uint64), ("rays_bounce", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("node_visits", C.c_uint64), ("leaf_prim_tests", C.c_uint64), ("subtree_visits", C.c_uint64),
                ("queue_pushes", C.c_uint64), ("staged_bytes", C.c_uint64), ("hit_updates", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("traverse_launches", C.c_uint64), ("ms_total", C.c_double),
                ("ms_traverse", C.c_double), ("ms_build", C.c_double), ("bvh_nodes", C.c_uint32),
                ("bvh_subtrees", C.c_uint32), ("bvh_levels", C.c_uint32), ("bvh_width", C.c_uint32),
                ("bvh_bytes", C.c_uint64), ("node_visits_l0", C.c_uint64), ("leaf_prim_tests_l0", C.c_uint64),
                ("queue_pushes_l0", C.c_uint64), ("staged_bytes_l0", C.c_uint64), ("hit_updates_l0", C.c_uint64),
                ("traverse_launches_l0", C.c_uint64), ("ms_traverse_l0", C.c_double),
                ("waves_retried", C.c_uint64), ("queues_grown", C.c_uint64), ("graph_replays", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}
