"""Generates tests/golden/ref_json_blobs.json: wavefunction documents encoded by the REFERENCE's own
pyANNonGPU/json_numpy.py NumpyEncoder (imported from /root/reference; only this module is importable here -- the Psi*.py
files need the compiled _pyANNonGPU), with exactly the keys the reference's to_json methods write
(pyANNonGPU/PsiRBM.py:7-19, PsiDeep.py:7-22, PsiCNN.py:7-23, PsiFullyPolarized.py:4-10).  The product must decode them
(tests/test_json_format.py).  Run in the build container: python tests/golden/make_json_fixture.py"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

spec = importlib.util.spec_from_file_location("ref_json_numpy", "/root/reference/pyANNonGPU/json_numpy.py")
ref_json = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_json)

fspec = importlib.util.spec_from_file_location("angpu_factories_standalone", os.path.join(ROOT, "annongpu_b200", "factories.py"))
F = importlib.util.module_from_spec(fspec)
sys.modules[fspec.name] = F
fspec.loader.exec_module(F)

rbm = F.rbm_spec(6, 12, noise=0.05, final_weight=3.0, seed=101)
deep = F.deep_spec(6, 6, [6, 3], [3, 2], noise=0.05, final_weights=2.0, seed=102)
cnn = F.cnn_spec([1, 2, 3], [(2, [1, 2, 2]), (1, [1, 2, 3])], noise=0.05, final_factor=1.5, seed=103)

docs = {
    "PsiRBM": dict(type="PsiRBM", W=rbm.W, final_weight=rbm.final_weight.real, log_prefactor_re=0.25, log_prefactor_im=-0.5),
    "PsiDeep": dict(type="PsiDeep", num_sites=deep.num_sites, a=deep.input_weights, b=list(deep.biases), connections=list(deep.connections),
                    W=list(deep.weights), final_weights=deep.final_weights, log_prefactor_re=0.1, log_prefactor_im=0.2),
    "PsiCNN": dict(type="PsiCNN", extent=np.asarray(cnn.extent, dtype=np.uint32), num_channels_list=np.asarray(cnn.num_channels_list, dtype=np.uint32),
                   connectivity_list=np.asarray(cnn.connectivity_list, dtype=np.uint32), symmetry_classes=np.asarray(cnn.symmetry_classes, dtype=np.uint32),
                   params=cnn.params, final_factor=cnn.final_factor, log_prefactor_re=-0.3, log_prefactor_im=0.0),
    "PsiFullyPolarized": dict(type="PsiFullyPolarized", num_sites=5, log_prefactor_re=-1.25, log_prefactor_im=0.0),
}
out = {k: json.loads(json.dumps(v, cls=ref_json.NumpyEncoder)) for k, v in docs.items()}
# round trip through the reference's decoder as a self-check of the fixture
back = json.loads(json.dumps(out["PsiRBM"]), cls=ref_json.NumpyDecoder)
assert np.array_equal(back["W"], rbm.W)
with open(os.path.join(HERE, "ref_json_blobs.json"), "w") as f:
    json.dump(out, f)
print("written", os.path.join(HERE, "ref_json_blobs.json"), os.path.getsize(os.path.join(HERE, "ref_json_blobs.json")), "bytes")
