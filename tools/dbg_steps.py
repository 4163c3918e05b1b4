import sys, ctypes as C, json
sys.path.insert(0,'/root/repo/tests')
from _util import *
lib=vp.load_product()
for idx in (2,3,4):
    sc,cam,fpd=scenes.config(idx)
    with vp.Session(lib) as s:
        s.set_scene(sc)
        lib.check(lib.c_stats_begin(s.ctx),"b")
        acc,_,st=s.render(cam, params(fpd, spp=1, flags=abi.VP_FLAG_COUNT_STEPS), want_rgb8=False)
        info=C.create_string_buffer(2048); lib.c_info(s.ctx, info, 2048); m=json.loads(info.value.decode())
        print(idx, st.as_dict(), "empty-region steps", m["dbg_extend_steps_in_empty_region"], "empty-brick steps", m["dbg_extend_steps_in_empty_brick"])
