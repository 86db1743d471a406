"""tfa.image.connected_components for the torch stand-in: the numpy stand-in's implementation on a torch tensor
(behind tf.stop_gradient in the reference, voting_layers_2d.py:37-56).  TEST INFRASTRUCTURE."""
import importlib.util as _u
import os as _os

import torch as _torch

_spec = _u.spec_from_file_location("_tfa_numpy_standin", _os.path.join(_os.path.dirname(__file__), "..", "..", "tf_standin",
                                                                        "tensorflow_addons", "__init__.py"))
_np_impl = _u.module_from_spec(_spec)
_spec.loader.exec_module(_np_impl)


class _Image:
    @staticmethod
    def connected_components(images, name=None):
        return _torch.as_tensor(_np_impl.image.connected_components(images.numpy()))


image = _Image()
