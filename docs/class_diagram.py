"""Architecture diagram generator (reference hivetrain/docs/test.py emits a 6-node graphviz diagram of TrainingLoop).
Writes docs/architecture.dot (render with `dot -Tpdf`); no graphviz python package needed."""
import os

EDGES = [
    ("neurons/miner.py", "training_manager.DeltaLoop"), ("neurons/validator.py", "validation_logic.DeltaValidator"),
    ("neurons/averager.py", "averaging_logic.ParameterizedAverager"),
    ("training_manager.DeltaLoop", "models.Trainer"), ("validation_logic.DeltaValidator", "models.Trainer"),
    ("averaging_logic.ParameterizedAverager", "models.Trainer"), ("models.Trainer", "models.TransformerEngine"),
    ("models.TransformerEngine", "ops (ctypes)"), ("ops (ctypes)", "csrc/*.cu -> libdtb200.so (sm_100a)"),
    ("ops (ctypes)", "ops.reference (PyTorch oracle, CPU)"),
    ("training_manager.DeltaLoop", "hf_manager.HFManager"), ("validation_logic.DeltaValidator", "hf_manager.HFManager"),
    ("averaging_logic.ParameterizedAverager", "hf_manager.HFManager"), ("hf_manager.HFManager", "parallel.exchange (peer|collective|disk)"),
    ("parallel.exchange (peer|collective|disk)", "parallel.symm.SymmetricWindow"), ("parallel.symm.SymmetricWindow", "csrc/symm_runtime.cu"),
    ("validation_logic.DeltaValidator", "btt_connector.BittensorNetwork"), ("averaging_logic.ParameterizedAverager", "chain_manager.ChainMultiAddressStore"),
    ("bench.py", "parallel.local_sgd.LocalSGDCoordinator"), ("parallel.local_sgd.LocalSGDCoordinator", "parallel.exchange (peer|collective|disk)"),
]
if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "architecture.dot"), "w") as f:
        f.write("digraph dtb200 {\n  rankdir=LR; node [shape=box, fontsize=10];\n")
        for a, b in EDGES:
            f.write(f'  "{a}" -> "{b}";\n')
        f.write("}\n")
    print("wrote docs/architecture.dot")
