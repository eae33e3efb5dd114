// wavefront kernels: see render_api.cu
