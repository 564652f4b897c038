"""Rank -> GPU mapping by PCIe topology (bandup_b200/numa.py): host logic, no GPU needed."""
from bandup_b200.numa import spread_order


def _hgx_paths():
    # two host bridges, two switches under each, two GPUs per switch: GPUs (0,1), (2,3), (4,5), (6,7)
    paths = []
    for g in range(8):
        hb, sw = g // 4, g // 2
        paths.append([f"pci0000:{hb:02x}", f"0000:{hb:02x}:01.0", f"0000:{10 + sw:02x}:00.0", f"0000:{20 + g:02x}:00.0"])
    return paths


def test_spread_order_takes_one_gpu_per_switch_first():
    order = spread_order(_hgx_paths())
    assert sorted(order) == list(range(8)) and order[0] == 0
    switches = lambda gpus: {g // 2 for g in gpus}                    # noqa: E731
    assert len(switches(order[:4])) == 4                                # four ranks: four different switches
    assert {g // 4 for g in order[:2]} == {0, 1}                        # two ranks: both host bridges


def test_spread_order_without_sysfs_is_bit_reversed():
    assert spread_order([[]] * 8) == [0, 4, 2, 6, 1, 5, 3, 7]
    assert sorted(spread_order([[]] * 6)) == list(range(6))
    assert spread_order([[]]) == [0] and spread_order([]) == []


def test_spread_order_uses_numa_nodes_when_the_tree_is_flat():
    # every GPU directly under its own root port: only the NUMA node tells neighbours apart
    paths = [[f"pci0000:{g:02x}", f"0000:{g:02x}:00.0"] for g in range(4)]
    order = spread_order(paths, [0, 0, 1, 1])
    assert {[0, 0, 1, 1][g] for g in order[:2]} == {0, 1}


def test_greedy_order_avoids_the_shared_host_path():
    """The 8-GPU boxes measured here: every GPU copies at 55 GB/s alone and in pairs, GPUs 0-3 share a
    115 GB/s path, GPUs 4-7 a 142 GB/s one, the host feeds 225 GB/s in all.  Four ranks must end up
    two and two, not on GPUs 0-3 (which is what index order gives: 29 GB/s each)."""
    from bandup_b200.numa import greedy_order

    def total(devs):
        a = min(115.0, 55.0 * sum(1 for d in devs if d < 4))
        b = min(142.0, 55.0 * sum(1 for d in devs if d >= 4))
        return min(225.0, a + b)
    order, prefix = greedy_order(8, total)
    assert sorted(order) == list(range(8)) and order[:2] == [0, 1]
    assert sum(1 for d in order[:4] if d < 4) == 2 and prefix[3] == 220.0
    assert total([0, 1, 2, 3]) == 115.0                      # what index order would have given
    # no sharing at all: index order
    assert greedy_order(4, lambda devs: 55.0 * len(devs))[0] == [0, 1, 2, 3]
    assert greedy_order(0, total) == ([], [])


def test_pinned_candidates_are_picked_by_rate_with_stable_ties():
    """bandup_b200.unfold.pick_fastest: the rule behind pinned_empty_near (the timing itself needs a GPU)."""
    from bandup_b200.unfold import pick_fastest
    assert pick_fastest([55.54, 55.55, 55.55, 55.53], 2) == [0, 1]          # all alike: first come
    assert pick_fastest([42.1, 55.5, 47.0, 55.4], 2) == [1, 3]              # two far blocks dropped
    assert pick_fastest([42.1, 55.5, 47.0, 43.0], 2) == [1, 2]              # one near block: next best by rate
    assert pick_fastest([50.0], 1) == [0]
