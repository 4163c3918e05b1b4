"""Frame sharding across ranks: interleaved row tiles.

Row y belongs to tile y // ROW_TILE; tile k is rendered by rank k % world.  The scene is
replicated, every rank accumulates only its own tiles across samples (no per-sample
communication) and the finished tiles are exchanged once per displayed frame with a single
all_gather of rank-contiguous packed tiles, then de-interleaved.  Works on any
torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

ROW_TILE = 8


def shard_params(world, rank):
    """VpFrameParams row-tile fields for this rank."""
    if world <= 1:
        return {}
    return dict(row_tile=ROW_TILE, row_tile_stride=world, row_tile_phase=rank)


def tiles_of_rank(height, world, rank):
    ntile = (height + ROW_TILE - 1) // ROW_TILE
    return list(range(rank, ntile, world))


class FrameGather:
    """Pre-allocated buffers for gathering a [H, W, C] (or [H, W]) frame whose rows are sharded."""

    def __init__(self, frame, world, rank):
        self.world, self.rank = world, rank
        self.H = frame.shape[0]
        assert self.H % ROW_TILE == 0, "sharded frames need a height that is a multiple of the row tile"
        self.ntile = self.H // ROW_TILE
        self.tail = tuple(frame.shape[1:])
        self.counts = [len(tiles_of_rank(self.H, world, r)) for r in range(world)]
        self.max_tiles = max(self.counts)
        self.pack = torch.zeros((self.max_tiles, ROW_TILE) + self.tail, dtype=frame.dtype, device=frame.device)
        self.out = torch.zeros((world, self.max_tiles, ROW_TILE) + self.tail, dtype=frame.dtype, device=frame.device)

    def __call__(self, frame):
        """frame: rows of this rank valid -> returns frame with every rank's rows filled in (in place)."""
        if self.world <= 1:
            return frame
        v = frame.view((self.ntile, ROW_TILE) + self.tail)
        self.pack[: self.counts[self.rank]] = v[self.rank::self.world]
        dist.all_gather_into_tensor(self.out.view(-1), self.pack.view(-1))
        # de-interleave: tile t of rank r is frame tile t * world + r.  The tile counts of the ranks differ by at most
        # one, so one permuted copy places every rank's first `full` tiles (this rank's own among them: same bytes) and
        # the ranks that own one more tile add it.
        full = min(self.counts)
        if full:
            v[: full * self.world].view((full, self.world, ROW_TILE) + self.tail).copy_(self.out[:, :full].transpose(0, 1))
        for r in range(self.world):
            if self.counts[r] > full and r != self.rank:
                v[full * self.world + r] = self.out[r, full]
        return frame
