"""save_particle_info (source/save_particle_info.cpp:21-130): the reference's on-disk dump, CPU-only format check.
Expected strings follow the default `std::ostream << float` format (printf "%g")."""
import os

import numpy as np

import apbf_b200


def _state():
    pos = np.zeros((4, 4), np.int32)
    # hidden slots; the index list picks 2, 0, 3 (slot 1 is not a fluid particle)
    pos[0, :3] = (0, 10 * 262144, -60 * 262144)            # at the centre: distance 0
    pos[2, :3] = (3 * 262144, 14 * 262144, -60 * 262144)   # distance 5
    pos[3, :3] = (131072, 10 * 262144, -60 * 262144 + 1)   # distance 0.5 (+ one unit in z: 0.5 in %g)
    return dict(index_list=np.array([2, 0, 3], np.uint32), position=pos, velocity=np.zeros((4, 4), np.float32),
                inverse_mass=np.array([0.125, 9.0, 1.0 / 3.0, 1e-5], np.float32), radius=np.array([1.0, 7.0, 1.25992107, 0.1], np.float32),
                pos_backup=pos.copy(), transferring=np.zeros(4, np.uint32), target_radius=np.array([1.5, 1.0, 123456.7], np.float32),
                kernel_width=np.array([4.0, 5.03968, 0.4], np.float32), boundariness=np.ones(3, np.float32),
                boundary_distance=np.array([262144, 393216, 0xFFFFFFFF], np.uint32))


def test_particle_info_files(tmp_path):
    pairs = np.array([[0, 1], [0, 2], [2, 0]], np.uint32)
    folder = apbf_b200.write_particle_info(_state(), pairs, str(tmp_path / "particle_data"))
    rd = lambda name: open(os.path.join(folder, name)).read()
    assert rd("centerDist.txt") == "5;0;0.5;"
    assert rd("radius.txt") == "1.25992;1;0.1;"                 # radius of the hidden slot behind each id
    assert rd("neighborCount.txt") == "2;0;1;"
    assert rd("kernelWidth.txt") == "4;5.03968;0.4;"
    assert rd("targetRadius.txt") == "1.5;1;123457;"
    assert rd("boundaryDistance.txt") == "1;1.5;16384;"
    lines = rd("data.csv").splitlines()
    assert lines[0] == "center distance,boundary distance,kernel width,neighbor count,radius,target radius,inverse mass,x,y,z"
    assert lines[1] == "0,1.5,5.03968,0,1,1,0.125,0,10,-60"          # sorted by the distance to (0, 10, -60)
    assert lines[2] == "0.5,16384,0.4,1,0.1,123457,1e-05,0.5,10,-60"
    assert lines[3] == "5,1,4,2,1.25992,1.5,0.333333,3,14,-60"
    assert len(lines) == 4
