"""Stand-in for torch_geometric==2.1.0.post1 (requirements.txt:8 of the reference).

TEST INFRASTRUCTURE ONLY.  The real package is not installable offline, so the four
entry points the reference's DGT hot path touches are restated in pure torch with the
PyG 2.1 semantics (see oracle/README.md).  Nothing under jodo_b200/ imports this.
"""
__version__ = "2.1.0.post1-shim"
