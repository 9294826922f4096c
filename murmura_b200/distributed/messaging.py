"""Wire format of the ZeroMQ compatibility backend.

Parity: reference ``murmura/distributed/messaging.py:25-78`` (``MsgType`` values, multipart frames,
``torch.save`` state payloads, pickle-4 objects).  Frame 0 is the reference's 5-byte ``"!Bi"`` header and frame 1
the payload, so an unmodified reference node or monitor decodes our messages.  One deliberate addition
(SURVEY §5.2): an OPTIONAL third frame carries the **round index** (``"!i"``) so a late MODEL_STATE from round k
can never be consumed in round k+1; the reference's ``decode`` ignores extra frames, ours reads it when present
(round = -1 = "unknown" otherwise; the round-1 9-byte ``"!Bii"`` header is still accepted on receive).

State payloads are loaded with ``weights_only=True`` (tensors only).  Metrics / topology claims stay pickle-4 for
wire parity with the reference; the PULL sockets bind to the node's configured host (``MURMURA_BIND_HOST`` overrides),
not to every interface, and must only be exposed on a trusted network, exactly as with the reference.
"""
from __future__ import annotations

import enum
import io
import pickle
import struct
from typing import Any, Dict, List, Tuple

import torch


class MsgType(enum.IntEnum):
    MODEL_STATE = 0
    METRICS = 1
    TOPO_CLAIM = 2


_HDR = struct.Struct("!Bi")
_HDR_R1 = struct.Struct("!Bii")                       # round-1 header of this backend (receive only)
_ROUND = struct.Struct("!i")
MONITOR_ID = -1


def encode(msg_type: MsgType, sender_id: int, payload: bytes, round_idx: int = -1) -> List[bytes]:
    frames = [_HDR.pack(int(msg_type), sender_id), payload]
    if round_idx >= 0:
        frames.append(_ROUND.pack(round_idx))
    return frames


def decode_full(frames: List[bytes]) -> Tuple[MsgType, int, int, bytes]:
    head, payload = frames[0], frames[1]
    rnd = -1
    if len(head) == _HDR_R1.size:
        kind, sender, rnd = _HDR_R1.unpack(head)
    else:
        kind, sender = _HDR.unpack(head)
    if len(frames) > 2 and len(frames[2]) == _ROUND.size:
        rnd = _ROUND.unpack(frames[2])[0]
    return MsgType(kind), sender, rnd, payload


def decode(frames: List[bytes]) -> Tuple[MsgType, int, bytes]:
    kind, sender, _, payload = decode_full(frames)
    return kind, sender, payload


def pack_state(state_dict: Dict[str, torch.Tensor]) -> bytes:
    buf = io.BytesIO()
    torch.save(state_dict, buf)
    return buf.getvalue()


def unpack_state(data: bytes) -> Dict[str, torch.Tensor]:
    return torch.load(io.BytesIO(data), map_location="cpu", weights_only=True)


def pack_obj(obj: Any) -> bytes:
    return pickle.dumps(obj, protocol=4)


def unpack_obj(data: bytes) -> Any:
    return pickle.loads(data)
