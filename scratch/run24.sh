for mb in 192 64 32 16 512; do PYTHONPATH=. FIMEX_B200_HOST_CHUNK_MB=$mb python scratch/e2e_probe.py; done
