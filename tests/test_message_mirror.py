"""The own-message shortcut (codegen.message_mirror / fgpu_launch.mirror): which agent variables a producer's message is
proved to carry unchanged -- decided from the script text, fail-closed -- and, on the GPU, that taking them from the message
changes nothing."""
import os

import numpy as np
import pytest

from common import REFERENCE, W, build
from ev_diffusion_b200 import codegen, xmml

BROWNIAN_XML = W.BROWNIAN_XML
HEAD = '#include "header.h"\n'
TAIL = ("\n__FLAME_GPU_FUNC__ int count_neighbours(xmachine_memory_Particle* agent, xmachine_message_location_list* location_messages, "
        "xmachine_message_location_PBM* partition_matrix){ return 0; }\n"
        "__FLAME_GPU_FUNC__ int move(xmachine_memory_Particle* agent, RNG_rand48* rand48){ return 0; }\n")


def _mirror(producer_body):
    model = xmml.parse_model(BROWNIAN_XML)
    src = HEAD + ("__FLAME_GPU_FUNC__ int output_location(xmachine_memory_Particle* agent, xmachine_message_location_list* location_messages)\n{"
                  + producer_body + "}\n") + TAIL
    consumer = next(f for f in model.all_functions() if f.name == "count_neighbours")
    return codegen.mirror_producer(model, consumer, [src])[1]


def test_plain_pass_through_is_mirrored():
    assert _mirror(" add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); return 0; ") == \
        {"id": "id", "x": "x", "y": "y", "z": "z"}
    shipped = open(W.brownian_function_files()[0]).read()
    model = xmml.parse_model(BROWNIAN_XML)
    consumer = next(f for f in model.all_functions() if f.name == "count_neighbours")
    producer, m = codegen.mirror_producer(model, consumer, [shipped])
    assert producer.name == "output_location" and m == {"id": "id", "x": "x", "y": "y", "z": "z"}
    assert codegen.mirror_producer(model, next(f for f in model.all_functions() if f.name == "move"), [shipped]) == (None, {})


@pytest.mark.parametrize("body,expected", [
    # a variable the producer writes (before or after the call) may differ from what the message holds
    (" add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); agent->x += 1.0f; return 0; ", {"id": "id", "y": "y", "z": "z"}),
    (" agent->z = 0.0f; add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); return 0; ", {"id": "id", "x": "x", "y": "y"}),
    # any other expression is not a pass-through
    (" add_location_message(location_messages, agent->id, agent->x + 0.5f, agent->y, -agent->z); return 0; ", {"id": "id", "y": "y"}),
    (" add_location_message(location_messages, agent->id, agent->y, agent->x, agent->z); return 0; ", {"id": "id", "y": "x", "x": "y", "z": "z"}),
    # the same variable twice: the first field is as good as the second; a constant is nothing
    (" add_location_message(location_messages, 7, agent->x, agent->x, agent->z); return 0; ", {"x": "x", "z": "z"}),
    # type mismatch (int variable into a float field, float into the int field)
    (" add_location_message(location_messages, agent->x, agent->id, agent->y, agent->z); return 0; ", {"y": "y", "z": "z"}),
    # conditional, repeated or unreachable output: nothing can be said about message k
    (" if (agent->id > 3) { add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); } return 0; ", {}),
    (" add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z);"
     " add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); return 0; ", {}),
    (" if (agent->nbr) return 0; add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); return 0; ", {}),
    (" for (int i = 0; i < 1; i++) { add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); } return 0; ", {}),
    # the agent pointer escapes: every variable counts as written
    (" helper(agent); add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); return 0; ", {}),
    # conditional compilation inside the body: what is compiled is not what the text shows
    ("\n#ifdef EMIT\n add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z);\n#endif\n return 0; ", {}),
    # comments and strings do not count
    (" /* add_location_message(a, b, c, d, e); */ add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); return 0; ",
     {"id": "id", "x": "x", "y": "y", "z": "z"}),
])
def test_fail_closed_cases(body, expected):
    assert _mirror(body) == expected


def test_a_macro_hiding_a_member_access_disables_it():
    model = xmml.parse_model(BROWNIAN_XML)
    consumer = next(f for f in model.all_functions() if f.name == "count_neighbours")
    src = (HEAD + "#define POS agent->x\n__FLAME_GPU_FUNC__ int output_location(xmachine_memory_Particle* agent, "
           "xmachine_message_location_list* location_messages){ POS = 1.0f; "
           "add_location_message(location_messages, agent->id, agent->x, agent->y, agent->z); return 0; }\n" + TAIL)
    assert codegen.mirror_producer(model, consumer, [src]) == (None, {})


@pytest.mark.skipif(not os.path.isdir(f"{REFERENCE}/examples"), reason="the reference tree is not mounted")
@pytest.mark.parametrize("name,function,expected", [
    ("Boids_Partitioning", "inputdata", {"id", "x", "y", "z", "fx", "fy", "fz"}),
    ("CirclesPartitioning_float", "inputdata", {"id"}),          # x, y, z go through locals: not provable from the call
    ("SmoothedParticleHydrodynamics", "computeDensityPressure", {"id", "x", "y", "z"}),
])
def test_reference_examples(name, function, expected):
    d = f"{REFERENCE}/examples/{name}/src/model"
    model = xmml.parse_model(f"{d}/XMLModelFile.xml")
    src = [open(f"{d}/{f}", errors="replace").read() for f in model.function_files]
    consumer = next(f for f in model.all_functions() if f.name == function)
    producer, m = codegen.mirror_producer(model, consumer, src)
    assert producer is not None and set(m) == expected, (producer, m)


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["c2_small", "c2_tiny"])
def test_results_do_not_depend_on_the_shortcut(key):
    """the same plugin with the shortcut on (default) and off: every variable and the seeds bit-identical after every iteration;
    the descriptor says the consumer may use it"""
    from common import brownian_plugin
    from ev_diffusion_b200.runtime import Simulation, load_model
    cfg = W.BROWNIAN[key]
    plugin = brownian_plugin(key, exact=True)
    model = load_model(plugin).contents
    names = [model.functions[i].name.decode() for i in range(model.nfunctions)]
    assert model.functions[names.index("count_neighbours")].mirror_producer == names.index("output_location")
    assert model.functions[names.index("move")].mirror_producer == -1
    on, off = Simulation(plugin), Simulation(plugin)
    off.set_option("mirror", 0)
    st = W.brownian_state(cfg, count=cfg.n - 37 if key == "c2_tiny" else None)
    for s in (on, off):
        s.set_agents("Particle", "default", st)
    for it in range(6):
        on.step(1)
        off.step(1)
        a, b = on.agents("Particle"), off.agents("Particle")
        for k in a:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (it, k)
        assert np.array_equal(on.seeds(), off.seeds())
    assert a["nbr"].max() > 0
    on.close()
    off.close()


def test_division_by_a_power_of_two_extent_is_an_exact_scaling():
    """fgpu::div_extent (csrc/fgpu_device.cuh) replaces the IEEE division of grid_position (krn:860-871) by a multiplication with
    the reciprocal when the extent of the partitioning bounds is a power of two.  Both are the correctly rounded value of the same
    real number, so they agree in every bit -- checked here in binary32 over values that include subnormal quotients, negative and
    huge numerators, infinities and NaN."""
    rng = np.random.default_rng(11)
    a = np.concatenate([
        rng.standard_normal(200000).astype(np.float32) * np.float32(1e3),
        (rng.random(200000).astype(np.float32) * np.float32(16384.0)),
        np.float32(2.0) ** rng.integers(-149, 127, 20000).astype(np.float32) * rng.standard_normal(20000).astype(np.float32),
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4028235e38, 1.17549435e-38, 16383.999], dtype=np.float32)])
    with np.errstate(all="ignore"):
        for ext in (np.float32(1.0), np.float32(8.0), np.float32(64.0), np.float32(128.0), np.float32(256.0), np.float32(0.5), np.float32(1024.0)):
            q = (a / ext).astype(np.float32)
            m = (a * (np.float32(1.0) / ext)).astype(np.float32)
            same = (q.view(np.uint32) == m.view(np.uint32)) | (np.isnan(q) & np.isnan(m))
            assert same.all(), (ext, a[~same][:5])
        # ... and it is NOT an identity for other extents, which is why those keep the division
        ext = np.float32(500.0)
        assert ((a / ext).astype(np.float32).view(np.uint32) != (a * (np.float32(1.0) / ext)).astype(np.float32).view(np.uint32)).any()
