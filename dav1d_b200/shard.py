"""Frame sharding over GPUs: frame n is reconstructed by rank n mod G — the device-side
analogue of dav1d's frame threads (`n_fc`, reference src/thread_task.c:409-436 for the dependency rule) — and
the only exchange on the data path is the finished reference picture, broadcast from its owner to every rank
that predicts from it. `torch.distributed` is plumbing: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests.
"""
import numpy as np


def frame_owner(n, world):
    return n % world


def decode_gop(frames, make_buffers, dist, rank, world, as_tensor, n_refs=2):
    """Reconstruct `frames` (list of synth frame dicts, decode order). Frame k predicts from the restored pictures
    of frames k-1 and k-2 (its own synthetic references stand in for pictures before the GOP).

    make_buffers(S) -> FrameBuffers on this rank's device; as_tensor(fb, name) -> the torch tensor aliasing one of
    its buffers (what broadcast sends / receives). Returns the list of restored pictures (numpy) on every rank."""
    pics = []          # per frame: tensor holding the restored picture on this rank
    keep = []
    for k, S in enumerate(frames):
        owner = frame_owner(k, world)
        fb = make_buffers(S)
        keep.append(fb)
        out = as_tensor(fb, fb.out_name)
        if rank == owner:
            # reference slots: most recent restored pictures first
            for slot in range(n_refs):
                if k - 1 - slot >= 0:
                    as_tensor(fb, "ref%d" % slot).copy_(pics[k - 1 - slot][:as_tensor(fb, "ref%d" % slot).numel()])
            fb.run()
            fb.alloc.sync()
        if world > 1:
            dist.broadcast(out, src=owner)
        pics.append(out)
    nbytes = frames[0]["pic"].nbytes
    return [p.cpu().numpy()[:nbytes].view(frames[0]["pic"].dtype).copy() for p in pics]
