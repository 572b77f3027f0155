import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
import b2rt
from b2rt.scene import Scene, place_camera
sc = Scene.load(os.path.join(ROOT, "scenes", "CBbunny.b2s")); cam = place_camera(sc, 1024, 768)
pt = b2rt.PathTracer(ns_aa=64, max_ray_depth=8, ns_area_light=1, seed=1)
pt.set_scene(sc); pt.set_camera(cam); pt.set_frame_size(1024, 768); pt.render()
for it in range(3):
    t = [time.perf_counter()]
    pt.set_scene(sc); t.append(time.perf_counter())
    pt.set_camera(cam); t.append(time.perf_counter())
    pt.start_raytracing(); t.append(time.perf_counter())
    pt.wait(); t.append(time.perf_counter())
    img = pt.rgba32f(); t.append(time.perf_counter())
    print("set_scene %.1f ms, set_camera %.1f, start %.1f, wait %.1f, read %.1f" % tuple((t[i+1]-t[i])*1e3 for i in range(5)), pt.stats()['ms_build'])
