from .segmentation.PyanNet import PyanNet
from .segmentation.PyanNet2 import PyanNet2
