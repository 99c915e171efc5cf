"""metamorph/train/train_mem.py:7-11 — launcher alias."""
from .train import train

if __name__ == "__main__":
    train()
