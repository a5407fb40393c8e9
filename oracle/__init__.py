"""CPU oracle of the retrieval-evaluation path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (hashgan_b200) never imports anything from here.
"""
