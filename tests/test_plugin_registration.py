"""Run-time MDP plugin registration (mctsb200_mdp_register), the CPU half: NVRTC compiles the tile kernel
around the plugin for sm_100a without a GPU, so the interface, the compiler-error path and the id bookkeeping are checked here;
the searches themselves are in tests/test_gpu_parity.py (-m gpu)."""
import ctypes as C

import pytest

import mcts_jl_b200 as m
from mcts_jl_b200 import abi


def _lib():
    return m.load_library()


def test_shipped_plugins_compile_and_get_stable_ids():
    mc = m.MountainCarPlugin()
    a = mc.mdp_id
    assert a >= abi.MDP_PLUGIN_BASE
    # the same definition again: same id, no recompilation needed
    assert m.register_mdp(mc.name, mc.source, mc.struct_name, 2, 3, 1) == a
    noisy = m.NoisyMountainCar()
    b = noisy.mdp_id
    assert b >= abi.MDP_PLUGIN_BASE and b != a
    # every host shape (2 / 4 state doubles x 2..4 actions) compiles
    ids = {m.NoisyMountainCar(actions=acts, state_width=sw).mdp_id for sw in (2, 4) for acts in ((-1.0, 1.0), (-1.0, 0.0, 1.0), (-1.0, -0.5, 0.5, 1.0))}
    assert len(ids) == 6 and b in ids
    # the InvertedPendulum of notebooks/Test_Visualization.ipynb cell 4 (det_sin, force noise) compiles too
    ip = m.InvertedPendulum()
    assert ip.mdp_id >= abi.MDP_PLUGIN_BASE and ip.mdp_id not in ids and ip.mdp_id != a
    lib = _lib()
    out = C.c_int32(-1)
    assert lib.mctsb200_mdp_lookup(mc.name.encode(), C.byref(out)) == abi.OK and out.value == a
    na, sb = C.c_int32(), C.c_int64()
    assert lib.mctsb200_mdp_info(b, C.byref(na), C.byref(sb)) == abi.OK
    assert (na.value, sb.value) == (3, 16)
    assert lib.mctsb200_mdp_info(abi.MDP_PLUGIN_BASE + 9999, C.byref(na), C.byref(sb)) == abi.ERR_INVALID_ARG


def test_compiler_errors_are_reported():
    with pytest.raises(m.MCTSB200Error) as e:
        m.register_mdp("Broken", "struct Broken { struct Params { double discount; }; };", "Broken", 2, 3, 1)
    assert e.value.status == abi.ERR_INVALID_ARG
    assert "does not compile" in e.value.message and "is_terminal" in e.value.message
    with pytest.raises(m.MCTSB200Error) as e:
        m.register_mdp("Syntax", "struct Syntax { this is not C++ };", "Syntax", 2, 3, 1)
    assert "Syntax.plugin" in e.value.message  # the log points into the user's text
    out = C.c_int32(-1)
    assert _lib().mctsb200_mdp_lookup(b"Broken", C.byref(out)) == abi.ERR_UNSUPPORTED  # a failed registration leaves nothing behind


def test_argument_checks():
    src = m.MountainCarPlugin().source
    for sd, na in [(3, 3), (2, 5), (2, 1), (8, 2)]:
        with pytest.raises(m.MCTSB200Error) as e:
            m.register_mdp("Shape", src, "MountainCarPlugin", sd, na, 1)
        assert e.value.status == abi.ERR_UNSUPPORTED
    with pytest.raises(m.MCTSB200Error) as e:
        m.register_mdp("MountainCar", src, "MountainCarPlugin", 2, 3, 1)  # built-in names stay built in
    assert e.value.status == abi.ERR_INVALID_ARG
    with pytest.raises(m.MCTSB200Error):
        m.register_mdp("Neg", src, "MountainCarPlugin", 2, 3, -1)
    # Params larger than the 256 bytes the ABI carries: rejected at compile time by the adapter's static_assert
    big = src.replace("double cost, jackpot;", "double cost, jackpot; double pad[40];")
    with pytest.raises(m.MCTSB200Error) as e:
        m.register_mdp("BigParams", big, "MountainCarPlugin", 2, 3, 1)
    assert "256 bytes" in e.value.message


def test_continuous_action_plugins_compile():
    """mctsb200_mdp_register_continuous: actions(mdp) = a Box (test/discussion_514.jl) for a model the caller brings. The DPW program
    compiles at registration, for the host shapes 2 / 4 state doubles x 1 ... 4 action doubles; the interface is checked there."""
    from mcts_jl_b200._lib import register_mdp_continuous

    cp = m.ContinuousPendulum()
    a = cp.mdp_id
    assert a >= abi.MDP_PLUGIN_BASE and cp.actions is None and cp.action_width == 1
    na, sb = C.c_int32(-1), C.c_int64()
    assert _lib().mctsb200_mdp_info(a, C.byref(na), C.byref(sb)) == abi.OK
    assert (na.value, sb.value) == (0, 16)  # no finite action set
    assert register_mdp_continuous(cp.name, cp.source, cp.struct_name, 2, 1) == a  # the same definition again
    # wider shapes: pad the state / the action (the dynamics ignore the padding)
    wide = cp.source.replace("        sp[1] = DM_ADD(w, DM_MUL(acc, p.dt));\n", "        sp[1] = DM_ADD(w, DM_MUL(acc, p.dt));\n        sp[2] = s[2];\n        sp[3] = s[3];\n")
    assert "sp[3] = s[3]" in wide
    two = cp.source.replace("rng.uniform01()));\n    }", "rng.uniform01()));\n        a[1] = rng.uniform01();\n    }", 1)
    assert "a[1] = rng.uniform01()" in two
    two = two.replace("        a[0] = u;\n", "        a[0] = u;\n        a[1] = 0.0;\n")
    ids = {register_mdp_continuous("CP41", wide, "ContinuousPendulum", 4, 1), register_mdp_continuous("CP22", two, "ContinuousPendulum", 2, 2),
           register_mdp_continuous("CP42", wide.replace("rng.uniform01()));\n    }", "rng.uniform01()));\n        a[1] = rng.uniform01();\n    }", 1)
                                   .replace("        a[0] = u;\n", "        a[0] = u;\n        a[1] = 0.0;\n"), "ContinuousPendulum", 4, 2)}
    assert len(ids) == 3 and a not in ids
    three = two.replace("        a[1] = rng.uniform01();\n", "        a[1] = rng.uniform01();\n        a[2] = rng.uniform01();\n").replace("        a[1] = 0.0;\n", "        a[1] = 0.0;\n        a[2] = 0.0;\n")
    assert register_mdp_continuous("CP23", three, "ContinuousPendulum", 2, 3) not in ids | {a}
    for sd, ad in [(2, 5), (3, 1), (2, 0), (8, 2)]:
        with pytest.raises(m.MCTSB200Error) as e:
            register_mdp_continuous("ShapeC", cp.source, "ContinuousPendulum", sd, ad)
        assert e.value.status == abi.ERR_UNSUPPORTED
    with pytest.raises(m.MCTSB200Error) as e:
        register_mdp_continuous("GaussianBox", cp.source, "ContinuousPendulum", 2, 1)
    assert e.value.status == abi.ERR_INVALID_ARG
    # a discrete plugin's struct does not have the continuous interface: the compiler says which member is missing
    with pytest.raises(m.MCTSB200Error) as e:
        register_mdp_continuous("NotContinuous", m.MountainCarPlugin().source, "MountainCarPlugin", 2, 1)
    assert "does not compile" in e.value.message and "next_action" in e.value.message


def test_redefinition_keeps_the_id():
    src = m.MountainCarPlugin().source
    a = m.register_mdp("Redefine", src, "MountainCarPlugin", 2, 3, 1)
    b = m.register_mdp("Redefine", src.replace("0.001", "0.002"), "MountainCarPlugin", 2, 3, 1)
    assert a == b


def test_plugin_policy_cfg():
    mdp = m.MountainCarPlugin()
    cfg, _ = m.build_cfg(m.DPWSolver(estimate_value=m.RolloutEstimator(m.PluginPolicy(100, (1, 2)))), mdp)
    assert cfg.rollout_policy == abi.POLICY_PLUGIN and list(cfg.rollout_policy_param)[:2] == [1, 2]
    with pytest.raises(ValueError):
        m.build_cfg(m.DPWSolver(estimate_value=m.RolloutEstimator(m.PluginPolicy(5))), mdp)


def test_compiled_programs_are_cached_on_disk(tmp_path):
    """A second process registers a known plugin from the on-disk cache ($MCTSB200_CACHE_DIR) without running the compiler; a
    changed definition misses."""
    import os
    import subprocess
    import sys

    code = ("import ctypes as C, sys, mcts_jl_b200 as m\n"
            "mdp = m.MountainCarPlugin(name='CacheProbe')\n"
            "if len(sys.argv) > 1: mdp.source = mdp.source.replace('0.001', '0.0015')\n"
            "mdp.mdp_id\n"
            "h, mi = C.c_int64(), C.c_int64()\n"
            "m.load_library().mctsb200_plugin_cache_stats(C.byref(h), C.byref(mi))\n"
            "print('STATS', h.value, mi.value)\n")
    env = dict(os.environ, MCTSB200_CACHE_DIR=str(tmp_path / "cache"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(*args):
        out = subprocess.run([sys.executable, "-c", code, *args], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        return [int(x) for x in out.stdout.split("STATS")[1].split()]

    assert run() == [0, 1]          # compiled, stored
    assert len(list((tmp_path / "cache").glob("plugin_*.bin"))) == 1
    assert run() == [1, 0]          # served from the cache
    assert run("changed") == [0, 1]  # another definition: another key
    env["MCTSB200_CACHE"] = "0"
    assert run() == [0, 1]          # cache off
