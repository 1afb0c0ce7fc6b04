"""Frame partitioning for multi-GPU runs: FFV1 with -g 1 is intra-only, so a
stream is split round-robin over the ranks (picture i -> rank i mod N) with no data-path
collective; packets are put back in presentation order on the host before muxing."""
import heapq


def frames_for_rank(nframes, rank, world):
    """indices of the pictures rank `rank` of `world` codes"""
    return list(range(rank, nframes, world))


def owner(frame_index, world):
    return frame_index % world


class ReorderQueue:
    """pts-ordered hand-over of packets coming back from several ranks / launch groups"""

    def __init__(self, first=0, step=1):
        self.next = first
        self.step = step
        self.heap = []

    def push(self, pts, item):
        heapq.heappush(self.heap, (pts, item))

    def pop_ready(self):
        out = []
        while self.heap and self.heap[0][0] == self.next:
            out.append(heapq.heappop(self.heap))
            self.next += self.step
        return out

    def __len__(self):
        return len(self.heap)
