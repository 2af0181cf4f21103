"""ctypes mirror of include/pt_b200.h (struct layouts + scene-description helpers).  No compute here."""
import ctypes as C
import json
import math
import os

SHAPES = {"SPHERE": 0, "CYLINDER": 1, "DISK": 2, "CONE": 3, "PARABOLOID": 4, "QUAD": 5, "CUBE": 6}
MATERIALS = {"LAMBERT": 0, "GGX": 1, "LAMBERT_GGX": 2}


class MaterialDesc(C.Structure):
    _fields_ = [("type", C.c_uint32), ("base_color", C.c_float * 3), ("emissive", C.c_float * 3),
                ("roughness", C.c_float), ("metalness", C.c_float), ("texture", C.c_uint32)]


class ObjectDesc(C.Structure):
    _fields_ = [("type", C.c_uint32), ("position", C.c_float * 3), ("rotation", C.c_float * 3),
                ("scale", C.c_float * 3), ("material", MaterialDesc)]


class CameraDesc(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("look_at", C.c_float * 3), ("up", C.c_float * 3),
                ("fovy", C.c_float), ("aspect", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64),
                ("shades", C.c_uint64), ("misses", C.c_uint64), ("gpu_ms", C.c_float), ("bvh_nodes", C.c_uint32),
                ("bvh_depth", C.c_uint32), ("scene_bytes", C.c_uint32), ("scene_in_smem", C.c_uint32)]


def f32(x):
    return C.c_float(x).value


def radians_f32(deg):
    """degree * (1.0f / 180.0f) * 3.14159265358979323846f in float arithmetic (reference SceneLoader.cpp:193-196)."""
    return f32(f32(f32(deg) * f32(1.0 / 180.0)) * f32(3.14159265358979323846))


def make_object(shape, position=(0, 0, 0), rotation_deg=(0, 0, 0), scale=(1, 1, 1), material="LAMBERT",
                base_color=(1, 1, 1), emissive=(0, 0, 0), roughness=0.5, metalness=0.0, texture=0):
    o = ObjectDesc()
    o.type = SHAPES[shape] if isinstance(shape, str) else int(shape)
    o.position[:] = [float(v) for v in position]
    o.rotation[:] = [radians_f32(v) for v in rotation_deg]
    o.scale[:] = [float(v) for v in scale]
    o.material.type = MATERIALS[material] if isinstance(material, str) else int(material)
    o.material.base_color[:] = [float(v) for v in base_color]
    o.material.emissive[:] = [float(v) for v in emissive]
    o.material.roughness = float(roughness)
    o.material.metalness = float(metalness)
    o.material.texture = int(texture)
    return o


def make_camera(position, look_at, fovy_deg, aspect, up=(0, 1, 0)):
    c = CameraDesc()
    c.position[:] = [float(v) for v in position]
    c.look_at[:] = [float(v) for v in look_at]
    c.up[:] = [float(v) for v in up]
    c.fovy = radians_f32(fovy_deg)
    c.aspect = float(aspect)
    return c


def object_array(objs):
    arr = (ObjectDesc * max(len(objs), 1))()
    for i, o in enumerate(objs):
        arr[i] = o
    return arr


def parse_scene_py(path, width, height):
    """Pure-Python restatement of the reference loadScene() schema (SceneLoader.cpp:124-348) used by the TESTS to
    cross-check the C++ loader (pt_parse_scene_file).  Returns (objects, texture_paths, skybox_path, camera)
    where each object's material.texture is an index+1 into texture_paths (0 = none).  Honours quirk Q5: scalar
    fields are only read from float literals."""
    with open(path) as f:
        text = f.read()
    # keep int/float distinction: python's json already parses 1 -> int, 1.0 -> float
    scene = json.loads(text)
    tex_paths = []

    def tex_handle(p):
        if not p:
            return 0
        if p in tex_paths:
            return tex_paths.index(p) + 1
        tex_paths.append(p)
        return len(tex_paths)

    def get_vec3(o, key, default):
        v = o.get(key)
        if isinstance(v, list) and len(v) == 3:
            return [float(x) for x in v]
        return list(default)

    def get_float(o, key, default):
        v = o.get(key)
        if isinstance(v, float):
            return v
        return default

    objs = []
    if isinstance(scene.get("objects"), list):
        for o in scene["objects"]:
            shape = o.get("type") if isinstance(o.get("type"), str) else "SPHERE"
            shape = shape if shape in SHAPES else "SPHERE"
            m = o.get("material") if isinstance(o.get("material"), dict) else None
            mt, bc, em, ro, me, tx = "LAMBERT", (1, 1, 1), (0, 0, 0), 0.5, 0.0, ""
            if m is not None:
                t = m.get("type")
                if isinstance(t, str) and t in MATERIALS:
                    mt = t
                bc = get_vec3(m, "baseColor", bc)
                em = get_vec3(m, "emissive", em)
                ro = get_float(m, "roughness", ro)
                me = get_float(m, "metalness", me)
                if isinstance(m.get("texture"), str):
                    tx = m["texture"]
            objs.append(make_object(shape, get_vec3(o, "position", (0, 0, 0)), get_vec3(o, "rotation", (0, 0, 0)),
                                    get_vec3(o, "scale", (1, 1, 1)), mt, bc, em, ro, me, tex_handle(tx)))
    sky = scene.get("skybox") if isinstance(scene.get("skybox"), str) else ""
    sky_idx = tex_handle(sky) if sky else 0
    cam = scene.get("camera") if isinstance(scene.get("camera"), dict) else {}
    camera = make_camera(get_vec3(cam, "position", (0, 0, 0)), get_vec3(cam, "look_at", (0, 0, -1)),
                         get_float(cam, "fovy", 60.0), f32(float(width)) / f32(float(height)))
    return objs, tex_paths, sky_idx, camera


REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(REPO_ROOT, "assets")
