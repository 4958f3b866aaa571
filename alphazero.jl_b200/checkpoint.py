"""Session checkpoint interop (src/ui/session.jl:64-118): the engine-side state of a training session
-- network weights and replay memory -- in a neutral binary format written NEXT TO the reference's Julia-`Serialization`
files (`bestnn.data`, `curnn.data`, `mem.data`, which only Julia can read), so that a Julia session can hand its state to
the engine and pick the engine's state up again (shim: `export_session` / `import_weights!` in julia/AlphaZeroB200.jl).

Files in a session directory (same directory layout as save_env, src/ui/session.jl:92-108):
  bestnn.azb, curnn.azb   network: b"AZB1" | u32 header length | JSON header | float32 parameter blob in Flux order
                          (the layout of az_net_load: Conv W[kw,kh,cin,cout] + b, BatchNorm gamma beta mu sigma2, Dense W[out,in] + b)
  mem.azs                 get_experience(env): b"AZS1" | u32 header length | JSON header | states u8[n*state_bytes] |
                          pi f64[n*A] (zero on illegal actions) | z f64[n] | t f64[n] | n i32[n]   (= az_samples_fetch / az_samples_from_host)
  iter.txt                env.itc as JSON, identical to the reference's file (session.jl:105-107)
  netparams.json          Network.hyperparams(bestnn) as JSON, identical to the reference's file (session.jl:99-101)
All integers little endian.  Pure host code (numpy); no GPU involved."""
import json
import os
import struct

import numpy as np

BESTNN_FILE, CURNN_FILE, MEM_FILE, ITC_FILE, NET_PARAMS_FILE = "bestnn.azb", "curnn.azb", "mem.azs", "iter.txt", "netparams.json"
_NET_MAGIC, _MEM_MAGIC = b"AZB1", b"AZS1"


def _write(path, magic, header, arrays):
    h = json.dumps(header, sort_keys=True).encode()
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(magic + struct.pack("<I", len(h)) + h)
        for a in arrays:
            f.write(np.ascontiguousarray(a).tobytes())
    os.replace(tmp, path)   # a reader never sees a half-written checkpoint


def _read(path, magic):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:4] != magic:
        raise ValueError("%s: not a %s file" % (path, magic.decode()))
    (n,) = struct.unpack("<I", raw[4:8])
    return json.loads(raw[8:8 + n].decode()), memoryview(raw)[8 + n:]


def hyperparams_dict(hp):
    """ResNetHP / SimpleNetHP -> the JSON object the reference writes to netparams.json (field names of the Julia structs)."""
    d = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(hp).items()}
    return d


def save_network(path, kind, game, hp, blob):
    """kind: "resnet" | "simplenet"; game: the name of src/examples.jl:17-21; hp: ResNetHP / SimpleNetHP (or a dict)."""
    blob = np.ascontiguousarray(blob, "<f4")
    hd = hp if isinstance(hp, dict) else hyperparams_dict(hp)
    _write(path, _NET_MAGIC, dict(kind=kind, game=game, hyperparams=hd, num_params=int(blob.size), dtype="float32", order="flux"), [blob])


def load_network(path):
    h, body = _read(path, _NET_MAGIC)
    blob = np.frombuffer(body, "<f4").copy()
    if blob.size != h["num_params"]:
        raise ValueError("%s: truncated (%d of %d parameters)" % (path, blob.size, h["num_params"]))
    return dict(kind=h["kind"], game=h["game"], hyperparams=h["hyperparams"], blob=blob)


def save_memory(path, game, state_bytes, num_actions, samples):
    """samples: dict(states, pi, z, t, n) as returned by Samples.fetch() (Vector{TrainingSample} in order)."""
    st = np.ascontiguousarray(samples["states"], np.uint8).reshape(-1, state_bytes)
    k = st.shape[0]
    pi = np.ascontiguousarray(samples["pi"], "<f8").reshape(k, num_actions)
    z, t, n = (np.ascontiguousarray(samples[x], d).reshape(k) for x, d in (("z", "<f8"), ("t", "<f8"), ("n", "<i4")))
    _write(path, _MEM_MAGIC, dict(game=game, num_samples=int(k), state_bytes=int(state_bytes), num_actions=int(num_actions)), [st, pi, z, t, n])


def load_memory(path):
    h, body = _read(path, _MEM_MAGIC)
    k, sb, A = h["num_samples"], h["state_bytes"], h["num_actions"]
    need = k * (sb + 8 * A + 8 + 8 + 4)
    if len(body) != need:
        raise ValueError("%s: expected %d payload bytes, found %d" % (path, need, len(body)))
    o = 0

    def take(dt, cnt, shape):
        nonlocal o
        a = np.frombuffer(body, dt, cnt, o).reshape(shape).copy()
        o += a.nbytes
        return a
    return dict(game=h["game"], states=take(np.uint8, k * sb, (k, sb)), pi=take("<f8", k * A, (k, A)), z=take("<f8", k, (k,)),
                t=take("<f8", k, (k,)), n=take("<i4", k, (k,)))


def valid_session_dir(d):   # src/ui/session.jl:84-90 for the neutral files
    return all(os.path.isfile(os.path.join(d, f)) for f in (BESTNN_FILE, CURNN_FILE, MEM_FILE, ITC_FILE))


def save_env(d, game, kind, hp, bestnn_blob, curnn_blob, state_bytes, num_actions, experience, itc):
    """save_env (src/ui/session.jl:92-108) for the engine-side state."""
    os.makedirs(d, exist_ok=True)
    save_network(os.path.join(d, BESTNN_FILE), kind, game, hp, bestnn_blob)
    save_network(os.path.join(d, CURNN_FILE), kind, game, hp, curnn_blob)
    save_memory(os.path.join(d, MEM_FILE), game, state_bytes, num_actions, experience)
    with open(os.path.join(d, NET_PARAMS_FILE), "w") as f:
        json.dump(hp if isinstance(hp, dict) else hyperparams_dict(hp), f, indent=2)
    with open(os.path.join(d, ITC_FILE), "w") as f:
        json.dump(int(itc), f)


def load_env(d):            # load_env (src/ui/session.jl:110-118)
    if not valid_session_dir(d):
        raise FileNotFoundError("%s is not a session directory with engine checkpoints" % d)
    with open(os.path.join(d, ITC_FILE)) as f:
        itc = json.load(f)
    return dict(bestnn=load_network(os.path.join(d, BESTNN_FILE)), curnn=load_network(os.path.join(d, CURNN_FILE)),
                experience=load_memory(os.path.join(d, MEM_FILE)), itc=itc)
