from .dino.model import DINOViT
