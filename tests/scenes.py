"""Scene inputs for the tests.

The reference's own corpus (Test/*.txt) is used where it has been staged under oracle/_ref/scenes by
`make -C oracle ref`; INLINE holds small scenes of our own in the same text format so that every test
also runs where the corpus is absent."""
import os

import oracle_py

INLINE = {
    "two_spheres_union": "Union\n Sphere -0.5 0 0 FF0000 1\n Sphere 0.5 0 0 00FF00 1\n",
    "two_spheres_diff": "Difference\n Sphere -0.5 0 0 FF0000 1\n Sphere 0.5 0 0 00FF00 1\n",
    "two_spheres_inter": "Intersection\n Sphere -0.5 0 0 FF0000 1\n Sphere 0.5 0 0 00FF00 1\n",
    "single_sphere": "Sphere 0 0 0 FF00FF 1.5",
    "single_cube": "Cube 0 0 0 12AB9F 2",
    "single_cylinder": "Cylinder 0 0 0 FF0056 1 3 30 40 50",
    # coincident faces: exact t ties decide pixels (Q5)
    "coincident_cubes": "Union\n Cube 0 0 -1 FF0000 2\n Cube 0 0 1 00FF00 2\n",
    "cube_minus_cylinder_caps": "Difference\n Cube 0 0 0 FF0000 2\n Cylinder 0 0 0 00FF00 0.7 2 0 0 0\n",
    "duplicate_spheres": "Union\n Sphere 0 0 0 FF0000 1\n Sphere 0 0 0 00FF00 1\n",
    # rotated long cylinder under an operator: the reference's non-conservative leaf box is observable (Q6)
    "rotated_cylinder_union": "Union\n Cylinder 0 0 0 00FF00 1 5 30 30 0\n Sphere 3 0 0 0000FF 0.5\n",
    "nested": ("Difference\n Intersection\n  Cube 0 0 0 FF0000 2\n  Sphere 0 0 0 0000FF 1.35\n Union\n  Union\n"
               "   Cylinder 0 0 0 00FF00 0.7 2.5 90 0 0\n   Cylinder 0 0 0 00FF00 0.7 2.5 0 0 0\n"
               "  Cylinder 0 0 0 00FF00 0.7 2.5 0 0 90\n"),
    "deep_left_chain": ("Union\n Union\n  Union\n   Union\n    Sphere -3 0 0 FF0000 1\n    Sphere -1.5 0 0 00FF00 1\n"
                        "   Cube 0 0 0 0000FF 1.5\n  Difference\n   Sphere 1.5 0 0 FFFF00 1\n   Sphere 1.9 0 0.5 FF00FF 0.6\n"
                        " Cylinder 3 0 0 00FFFF 0.5 2 45 0 45\n"),
}


def _sphere_chain(seed, n, wrap="Difference\nCube 0 0 0 FFD000 5\n"):
    """`wrap` around a balanced Union of n overlapping spheres scattered along and around the view axis: long tunnels through
    the cube, many run extensions — the inputs of the flat evaluation of sphere unions (flat_eval, csg_kernel.cuh)."""
    import random
    rnd = random.Random(seed)
    leaves = [f"Sphere {rnd.uniform(-1.2, 1.2):.4f} {rnd.uniform(-1.2, 1.2):.4f} {rnd.uniform(-3.0, 3.0):.4f} 30A0F0 {rnd.uniform(0.3, 0.9):.4f}"
              for _ in range(n)]

    def union(xs):
        if len(xs) == 1:
            return xs[0]
        h = len(xs) // 2
        return "Union\n" + union(xs[:h]) + "\n" + union(xs[h:])
    return wrap + union(leaves) + "\n"


INLINE.update({
    "sphere_chain_12": _sphere_chain(1, 12),                    # one simple flat Union under a Difference
    "sphere_chain_26": _sphere_chain(2, 26),                    # a composite flat Union (2 x 13)
    "sphere_chain_40": _sphere_chain(3, 40),                    # flat Unions below ordinary ones
    "sphere_union_root": _sphere_chain(4, 14, wrap=""),         # the whole tree is one flat Union
    "cube_and_spheres": _sphere_chain(5, 18, wrap="Intersection\nCube 0 0 0 FFD000 3\n"),
    "spheres_minus_sphere": "Difference\n" + _sphere_chain(6, 10, wrap="") + "Sphere 0 0 0 FF0000 1.1\n",   # flat left operand
})


# duplicated and concentric spheres inside a chain: exact ties between roots of different primitives (flat_eval gives up).  Not
# in INLINE: with exact ties the re-balanced tree (optimize = 1) is not the reference's tree; used with optimize = 0.
DUP_CHAIN = ("Difference\nCube 0 0 0 FFD000 5\nUnion\nUnion\nSphere 0 0 2 FF0000 0.8\nSphere 0 0 2 00FF00 0.8\nUnion\nUnion\n"
             "Sphere 0 0 1 0000FF 0.8\nSphere 0 0 1 0000FF 0.5\nUnion\nSphere 0.1 0 0.2 FF00FF 0.8\nSphere 0.1 0 0.2 FFFF00 0.8\n")


def corpus_names():
    d = oracle_py.SCENES_DIR
    if not os.path.isdir(d):
        return []
    return sorted(fn[:-4] for fn in os.listdir(d) if fn.endswith(".txt"))


def all_scene_ids():
    return ["inline:" + k for k in INLINE] + ["corpus:" + k for k in corpus_names()]


def text_of(scene_id):
    kind, name = scene_id.split(":", 1)
    if kind == "inline":
        return INLINE[name].encode()
    return oracle_py.scene_text(name)
