# Build recipe for pixelforge-b200 (also driven by __graft_entry__.build()).
#
#   make lib      pixelforge_b200/lib/libpixelforge.so       product: C99 front end + sm_100a kernels
#   make scenes   pixelforge_b200/lib/libpfscenes_cuda.so    benchmark/parity scenes on the product
#   make oracle   oracle/_build/libpixelforge_oracle.so      TEST ONLY: front end + scalar C restatement
#                 oracle/_build/libpfscenes_oracle.so
#   make ref      oracle/_ref/libpf_ref*.so, libpfscenes_ref*.so   TEST ONLY: the unmodified reference,
#                 compiled from /root/reference where it lies (skipped when it is absent)
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        := gcc
# -ffp-contract=off / -fmad=false: the reference is built without FMA (CMakeLists.txt:45-48); a fused
# multiply-add anywhere in the vertex or fragment arithmetic would change pixels.
HOST_CFLAGS = -std=gnu99 -O2 -fPIC -ffp-contract=off -fno-fast-math -fvisibility=hidden -DPF_BUILD_SHARED -DNDEBUG \
              -Iinclude -Wall -Wextra -Wno-unused-parameter -Wno-missing-field-initializers
NVCC_FLAGS  = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Ipixelforge_b200/csrc \
              -Xcompiler -fPIC,-fvisibility=hidden -Iinclude

HOST_SRC = pixelforge_b200/csrc/host/pf_context.c pixelforge_b200/csrc/host/pf_pipeline.c \
           pixelforge_b200/csrc/host/pf_objects.c pixelforge_b200/csrc/host/pf_x86approx.c
HOST_HDR = include/pixelforge.h include/pfcu.h include/pfx.h pixelforge_b200/csrc/host/pf_internal.h pixelforge_b200/csrc/host/pf_math.h \
           pixelforge_b200/csrc/pf_vstage.h pixelforge_b200/csrc/pf_prims.h pixelforge_b200/csrc/pf_pixfmt.h
HOST_OBJ = $(patsubst pixelforge_b200/csrc/host/%.c,build/host/%.o,$(HOST_SRC))
SCENES   = pixelforge_b200/scenes/scenes.c
LIBDIR   = pixelforge_b200/lib

all: lib scenes oracle ref

lib: $(LIBDIR)/libpixelforge.so
scenes: $(LIBDIR)/libpfscenes_cuda.so
oracle: oracle/_build/libpixelforge_oracle.so oracle/_build/libpfscenes_oracle.so
ref:
	bash oracle/build_ref.sh
	@if [ -d /root/reference/src ]; then $(MAKE) --no-print-directory refscenes; fi
refscenes: oracle/_ref/libpfscenes_ref.so oracle/_ref/libpfscenes_ref_bfix.so

build/host/%.o: pixelforge_b200/csrc/host/%.c $(HOST_HDR)
	@mkdir -p build/host
	$(CC) $(HOST_CFLAGS) -c $< -o $@

build/pfcu.o: pixelforge_b200/csrc/pfcu.cu $(wildcard pixelforge_b200/csrc/*.cuh) include/pfcu.h pixelforge_b200/csrc/pf_vstage.h pixelforge_b200/csrc/pf_prims.h pixelforge_b200/csrc/pf_pixfmt.h
	@mkdir -p build
	$(NVCC) $(NVCC_FLAGS) $(NVCC_EXTRA) -Xptxas -v -c $< -o $@ 2> build/pfcu.ptxas.log || (cat build/pfcu.ptxas.log; false)

$(LIBDIR)/libpixelforge.so: $(HOST_OBJ) build/pfcu.o
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared -o $@ $(HOST_OBJ) build/pfcu.o -Xlinker -soname,libpixelforge.so

$(LIBDIR)/libpfscenes_cuda.so: $(SCENES) $(LIBDIR)/libpixelforge.so include/pixelforge.h include/pfx.h
	$(CC) -std=gnu99 -O2 -fPIC -shared -Iinclude -DPFSCENE_HAVE_PFX -o $@ $(SCENES) -L$(LIBDIR) -lpixelforge -lm \
	    -Wl,-rpath,'$$ORIGIN'

build/pfcu_oracle.o: oracle/pfcu_oracle.c include/pfcu.h pixelforge_b200/csrc/pf_vstage.h pixelforge_b200/csrc/pf_prims.h pixelforge_b200/csrc/pf_pixfmt.h
	@mkdir -p build
	$(CC) -std=gnu99 -O2 -fPIC -ffp-contract=off -fno-fast-math -msse4.1 -Iinclude -c $< -o $@

oracle/_build/libpixelforge_oracle.so: $(HOST_OBJ) build/pfcu_oracle.o
	@mkdir -p oracle/_build
	$(CC) -shared -o $@ $(HOST_OBJ) build/pfcu_oracle.o -lm -Wl,-soname,libpixelforge_oracle.so

oracle/_build/libpfscenes_oracle.so: $(SCENES) oracle/_build/libpixelforge_oracle.so
	$(CC) -std=gnu99 -O2 -fPIC -shared -Iinclude -DPFSCENE_HAVE_PFX -o $@ $(SCENES) -Loracle/_build -lpixelforge_oracle -lm \
	    -Wl,-rpath,'$$ORIGIN'

# the same scene code against the unmodified reference (its own header, its own library)
oracle/_ref/libpfscenes_ref.so: $(SCENES) oracle/_ref/libpf_ref.so
	$(CC) -std=gnu99 -O2 -fPIC -shared -I/root/reference/src -fopenmp -o $@ $(SCENES) -Loracle/_ref -lpf_ref -lm -Wl,-rpath,'$$ORIGIN'
oracle/_ref/libpfscenes_ref_bfix.so: $(SCENES) oracle/_ref/libpf_ref_bfix.so
	$(CC) -std=gnu99 -O2 -fPIC -shared -I/root/reference/src -fopenmp -o $@ $(SCENES) -Loracle/_ref -lpf_ref_bfix -lm -Wl,-rpath,'$$ORIGIN'

clean:
	rm -rf build $(LIBDIR) oracle/_build

.PHONY: all lib scenes oracle ref refscenes clean
