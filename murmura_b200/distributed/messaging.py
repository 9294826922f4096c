"""Wire format of the ZeroMQ compatibility backend.

Parity: reference ``murmura/distributed/messaging.py:25-78`` (``MsgType`` values, two-frame
multipart, ``torch.save`` state payloads, pickle-4 objects).  One deliberate fix
(SURVEY §5.2): the header carries the **round index** (``"!Bii"`` instead of ``"!Bi"``) so a late
MODEL_STATE from round k can never be consumed in round k+1; ``decode`` still accepts the
reference's 5-byte header (round = -1 = "unknown").
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


_HDR = struct.Struct("!Bii")
_HDR_LEGACY = struct.Struct("!Bi")
MONITOR_ID = -1


def encode(msg_type: MsgType, sender_id: int, payload: bytes, round_idx: int = -1) -> List[bytes]:
    return [_HDR.pack(int(msg_type), sender_id, round_idx), payload]


def decode_full(frames: List[bytes]) -> Tuple[MsgType, int, int, bytes]:
    head, payload = frames[0], frames[1]
    if len(head) == _HDR.size:
        kind, sender, rnd = _HDR.unpack(head)
    else:
        kind, sender = _HDR_LEGACY.unpack(head)
        rnd = -1
    return MsgType(kind), sender, rnd, payload


def decode(frames: List[bytes]) -> Tuple[MsgType, int, bytes]:
    kind, sender, _, payload = decode_full(frames)
    return kind, sender, payload


def pack_state(state_dict: Dict[str, torch.Tensor]) -> bytes:
    buf = io.BytesIO()
    torch.save(state_dict, buf)
    return buf.getvalue()


def unpack_state(data: bytes) -> Dict[str, torch.Tensor]:
    return torch.load(io.BytesIO(data), map_location="cpu", weights_only=False)


def pack_obj(obj: Any) -> bytes:
    return pickle.dumps(obj, protocol=4)


def unpack_obj(data: bytes) -> Any:
    return pickle.loads(data)
