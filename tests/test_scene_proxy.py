"""SceneProxy (renderer/src/concurrent/scene_proxy.rs): the scene on a worker thread. CPU tests: the command stream a
proxy hands over equals the one a direct build sends; messages are applied in order; errors and teardown."""
import ctypes as C
import threading

import numpy as np
import pytest

from pathfinder_b200 import _lib as L
from pathfinder_b200 import api, scenes


def digest(cmd):
    """What identifies a command: its kind and, for the two big ones, the payload bytes."""
    kind = int(cmd.kind)
    if kind == L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11:
        up = cmd.u.upload_scene_d3d11
        d = up.draw_segments
        pts = np.ctypeslib.as_array(C.cast(d.points, C.POINTER(C.c_float)), shape=(int(d.point_count) * 2,)).copy() if d.point_count else np.zeros(0, np.float32)
        return (kind, int(d.point_count), int(d.index_count), pts.tobytes())
    if kind == L.PF_RENDER_COMMAND_DRAW_TILES_D3D11:
        b = cmd.u.draw_tiles_d3d11.tile_batch_data
        n = int(b.path_count)
        infos = C.cast(b.prepare_info.tile_path_info, C.POINTER(L.PFTilePathInfoD3D11))
        rects = tuple((infos[i].tile_min_x, infos[i].tile_min_y, infos[i].tile_max_x, infos[i].tile_max_y, infos[i].color) for i in range(n))
        return (kind, n, int(b.tile_count), int(b.segment_count), rects)
    return (kind,)


def direct_stream(flat, options=None, view_box=None):
    scene = api.Scene.from_flat(flat)
    if view_box is not None:
        scene.set_view_box(view_box)
    out = []
    scene.build(options or api.BuildOptions(), lambda cmd: out.append(digest(cmd)))
    return out


def small_scene(seed=1, n=300, size=512):
    return scenes.random_paths(n, size, seed)


def test_proxy_stream_equals_direct_build():
    flat = small_scene()
    want = direct_stream(flat)
    proxy = api.SceneProxy(api.Scene.from_flat(flat))
    got = []
    proxy.build(api.BuildOptions())
    proxy.receive(lambda cmd: got.append(digest(cmd)))
    assert got == want
    assert got[0][0] == L.PF_RENDER_COMMAND_START and got[-1][0] == L.PF_RENDER_COMMAND_FINISH
    # a second build of the unchanged scene: the proxy's own sink remembers the scene (SceneSink.last_scene)
    again = []
    proxy.build(api.BuildOptions())
    proxy.receive(lambda cmd: again.append(digest(cmd)))
    assert [d[0] for d in again] == [d[0] for d in want if d[0] != L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11]
    proxy.close()


def test_messages_are_applied_in_order_and_builds_queue():
    flat = small_scene(seed=2)
    proxy = api.SceneProxy(api.Scene.from_flat(flat))
    half = (0.0, 0.0, 256.0, 256.0)
    t = api.Transform2F(0.5, 0.0, 0.0, 0.5, 3.0, 4.0)
    proxy.build(api.BuildOptions())                # build 1: the scene as it is
    proxy.set_view_box(half)
    proxy.build(api.BuildOptions(transform=t))     # build 2: smaller view box, a transform
    other = small_scene(seed=3, n=50)
    proxy.replace_scene(api.Scene.from_flat(other))
    proxy.build(api.BuildOptions())                # build 3: another scene
    streams = []
    for _ in range(3):
        got = []
        proxy.receive(lambda cmd: got.append(digest(cmd)))
        streams.append(got)
    assert streams[0] == direct_stream(flat)
    assert streams[1] == direct_stream(flat, api.BuildOptions(transform=t), view_box=half)
    assert streams[2] == direct_stream(other)
    with pytest.raises(L.PathfinderCudaError) as e:  # nothing left to render
        proxy.receive(lambda cmd: None)
    assert e.value.status == L.PF_CUDA_ERROR_PROTOCOL
    proxy.close()


def test_build_runs_beside_the_caller():
    """build() returns before the build is done; the worker gets as far as its first command by itself."""
    flat = small_scene(seed=4, n=3000, size=2048)
    proxy = api.SceneProxy(api.Scene.from_flat(flat))
    proxy.build(api.BuildOptions())
    seen = []
    done = threading.Event()

    def pump():
        proxy.receive(lambda cmd: seen.append(int(cmd.kind)))
        done.set()

    th = threading.Thread(target=pump)
    th.start()
    assert done.wait(30.0)
    th.join()
    assert seen[-1] == L.PF_RENDER_COMMAND_FINISH
    proxy.close()


def test_copy_scene_and_listener_errors():
    flat = small_scene(seed=5, n=40)
    proxy = api.SceneProxy(api.Scene.from_flat(flat))
    proxy.set_view_box((0.0, 0.0, 300.0, 200.0))
    copy = proxy.copy_scene()  # waits for the set_view_box before it
    assert copy.view_box() == (0.0, 0.0, 300.0, 200.0)
    got = []
    copy.build(api.BuildOptions(), lambda cmd: got.append(digest(cmd)))
    assert got == direct_stream(flat, view_box=(0.0, 0.0, 300.0, 200.0))
    # a listener that fails aborts the build and the status comes back
    proxy.build(api.BuildOptions())

    def bad(cmd):
        raise L.PathfinderCudaError(L.PF_CUDA_ERROR_UNSUPPORTED, "no")

    with pytest.raises(L.PathfinderCudaError) as e:
        proxy.receive(bad)
    assert e.value.status == L.PF_CUDA_ERROR_UNSUPPORTED
    # the proxy is usable afterwards
    ok = []
    proxy.build(api.BuildOptions())
    proxy.receive(lambda cmd: ok.append(int(cmd.kind)))
    assert ok[-1] == L.PF_RENDER_COMMAND_FINISH
    # copy_scene with a build still to be rendered would never be reached: refused
    proxy.build(api.BuildOptions())
    with pytest.raises(L.PathfinderCudaError):
        proxy.copy_scene()
    proxy.close()  # tears down with a build in flight


@pytest.mark.gpu
def test_proxy_renders_the_same_frame():
    flat = scenes.random_paths(2000, 1024, 7)
    r = api.CudaRenderer((1024, 1024), background_color=(1, 1, 1, 1))
    api.Scene.from_flat(flat).build_and_render(r, api.BuildOptions())
    want = r.read_pixels()
    r.close()
    r = api.CudaRenderer((1024, 1024), background_color=(1, 1, 1, 1))
    proxy = api.SceneProxy(api.Scene.from_flat(flat))
    for _ in range(3):  # (first frame sizes the stages, the next ones take the steady-state path)
        proxy.build_and_render(r, api.BuildOptions())
    assert np.array_equal(r.read_pixels(), want)
    # a new view box, then build + render as two calls
    proxy.set_view_box((0.0, 0.0, 512.0, 512.0))
    proxy.build(api.BuildOptions())
    proxy.render(r)
    got = r.read_pixels()
    r.close()
    r = api.CudaRenderer((1024, 1024), background_color=(1, 1, 1, 1))
    s = api.Scene.from_flat(flat)
    s.set_view_box((0.0, 0.0, 512.0, 512.0))
    s.build_and_render(r, api.BuildOptions())
    assert np.array_equal(got, r.read_pixels())
    r.close()
    proxy.close()


def test_scene_clone_is_a_deep_copy_with_the_same_identity():
    """Scene: Clone (scene.rs:37) keeps id and epoch — a sink that has seen the original takes the clone for the same
    scene — and shares nothing with it afterwards."""
    flat = small_scene(seed=8, n=60)
    a = api.Scene.from_flat(flat)
    h = L.lib().PFSceneClone(a._h)
    b = api.Scene.__new__(api.Scene)
    b._h = h
    assert L.lib().PFSceneGetEpoch(a._h) == L.lib().PFSceneGetEpoch(b._h)
    sink = L.PFSceneSinkState()
    first, second = [], []
    a.build(api.BuildOptions(), lambda cmd: first.append(digest(cmd)), sink_state=sink)
    b.build(api.BuildOptions(), lambda cmd: second.append(digest(cmd)), sink_state=sink)
    assert L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11 in [d[0] for d in first]
    assert L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11 not in [d[0] for d in second]  # same scene, same epoch: not re-sent
    assert [d for d in second] == [d for d in first if d[0] != L.PF_RENDER_COMMAND_UPLOAD_SCENE_D3D11]
    # the original changes; the clone does not
    a.set_view_box((0.0, 0.0, 100.0, 100.0))
    assert b.view_box() == tuple(float(v) for v in flat.view_box)
    third = []
    b.build(api.BuildOptions(), lambda cmd: third.append(digest(cmd)), sink_state=L.PFSceneSinkState())
    assert third == first
