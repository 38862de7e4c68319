"""ctypes mirror of include/pbrgpu.h (POD descriptors only; no logic)."""
import ctypes as C

MAT_MATTE, MAT_MIRROR, MAT_GLASS, MAT_METAL, MAT_PLASTIC = range(5)
LIGHT_POINT, LIGHT_AREA_TRI, LIGHT_INFINITE, LIGHT_SKYBOX, LIGHT_AREA_SPHERE = range(5)
CAMERA_PERSPECTIVE, CAMERA_ORTHOGRAPHIC = 0, 1
SAMPLER_HALTON, SAMPLER_SOBOL = 0, 1
INTEGRATOR_PATH, INTEGRATOR_DIRECT_ONE, INTEGRATOR_DIRECT_ALL, INTEGRATOR_WHITTED, INTEGRATOR_VOLPATH = range(5)
MEDIUM_HOMOGENEOUS = 0
LIGHTSTRATEGY_UNIFORM, LIGHTSTRATEGY_POWER = 0, 1
RENDER_ACCUMULATE, RENDER_COUNTERS, RENDER_TIMERS, RENDER_LOCAL_ONLY = 1, 2, 4, 8
DUMP_SAMPLER, DUMP_PRIMARY_RAYS, DUMP_PRIMARY_HITS, DUMP_PERMS, DUMP_SHADOW, DUMP_HIT_DIFFERENTIALS, DUMP_WAVE_HITS = range(7)
TRACE_VIA_WAVE_KERNEL = 2

f3 = C.c_float * 3
f16 = C.c_float * 16


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("remap_roughness", C.c_int32), ("sigma", C.c_float), ("u_rough", C.c_float),
                ("kd", f3), ("v_rough", C.c_float), ("kr", f3), ("ks_texture_plus1", C.c_int32), ("kt", f3), ("kd_texture_plus1", C.c_int32),
                ("eta", f3), ("rough_texture_plus1", C.c_int32), ("k", f3), ("pad3", C.c_float)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("tri", C.c_int32), ("two_sided", C.c_int32), ("n_samples", C.c_int32),
                ("pos", f3), ("world_radius", C.c_float), ("L", f3), ("pad0", C.c_float),
                ("light_to_world", f16), ("world_to_light", f16)]


WRAP_REPEAT, WRAP_BLACK, WRAP_CLAMP = range(3)


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("texels", C.POINTER(C.c_float)), ("do_trilinear", C.c_int32),
                ("max_anisotropy", C.c_float), ("wrap", C.c_int32), ("su", C.c_float), ("sv", C.c_float), ("du", C.c_float),
                ("dv", C.c_float), ("pad0", C.c_int32)]


class Medium(C.Structure):
    _fields_ = [("type", C.c_int32), ("g", C.c_float), ("sigma_a", f3), ("sigma_s", f3)]


class Sphere(C.Structure):
    _fields_ = [("center", f3), ("radius", C.c_float), ("material", C.c_int32), ("light", C.c_int32),
                ("reverse_orientation", C.c_int32), ("pad0", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("n_verts", C.c_int32), ("n_tris", C.c_int32),
                ("verts", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_int32)),
                ("normals", C.POINTER(C.c_float)), ("uvs", C.POINTER(C.c_float)),
                ("tri_material", C.POINTER(C.c_int32)), ("tri_light", C.POINTER(C.c_int32)),
                ("tri_flags", C.POINTER(C.c_uint8)),
                ("n_materials", C.c_int32), ("n_lights", C.c_int32),
                ("materials", C.POINTER(Material)), ("lights", C.POINTER(Light)),
                ("env_width", C.c_int32), ("env_height", C.c_int32), ("env_texels", C.POINTER(C.c_float)),
                ("n_media", C.c_int32), ("pad0", C.c_int32), ("media", C.POINTER(Medium)),
                ("tri_medium_inside", C.POINTER(C.c_int32)), ("tri_medium_outside", C.POINTER(C.c_int32)),
                ("n_textures", C.c_int32), ("pad1", C.c_int32), ("textures", C.POINTER(Texture)),
                ("n_spheres", C.c_int32), ("pad2", C.c_int32), ("spheres", C.POINTER(Sphere)),
                ("tangents", C.POINTER(C.c_float))]


class CameraDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("lens_radius", C.c_float), ("focal_distance", C.c_float),
                ("raster_to_camera", f16), ("camera_to_world", f16), ("medium_plus1", C.c_int32)]


class SamplerDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("sample_at_pixel_center", C.c_int32), ("samples_per_pixel", C.c_int64),
                ("sample_bounds", C.c_int32 * 4), ("perms", C.POINTER(C.c_uint16)), ("n_perms", C.c_int64),
                ("sobol_matrices32", C.POINTER(C.c_uint32)), ("sobol_n_dims", C.c_int32), ("pad0", C.c_int32),
                ("sobol_vdc", C.POINTER(C.c_uint64)), ("sobol_vdc_inv", C.POINTER(C.c_uint64))]


class RenderDesc(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("max_depth", C.c_int32), ("rr_threshold", C.c_float),
                ("light_strategy", C.c_int32), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32),
                ("y1", C.c_int32), ("s0", C.c_int64), ("s1", C.c_int64), ("samples_per_wave", C.c_int32),
                ("flags", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("camera_samples", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_total", C.c_double), ("ms_raygen", C.c_double), ("ms_trace_closest", C.c_double),
                ("ms_shade", C.c_double), ("ms_trace_shadow", C.c_double), ("ms_trace_mis", C.c_double),
                ("ms_film", C.c_double), ("n_trace_closest_launches", C.c_uint64), ("rays_path", C.c_uint64)]


class DumpDesc(C.Structure):
    _fields_ = [("what", C.c_int32), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("s0", C.c_int64), ("s1", C.c_int64), ("n_dims", C.c_int32), ("pad", C.c_int32)]


assert C.sizeof(Sphere) == 32 and C.sizeof(Material) == 96 and C.sizeof(Light) == 176 and C.sizeof(Medium) == 32 and C.sizeof(Texture) == 48
