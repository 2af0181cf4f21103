/* Empty stand-in for <GL/gl.h>.  The CUDA toolkit's cuda_gl_interop.h includes it, and the reference's
 * Pathtracer.cpp includes cuda_gl_interop.h; the headless path never touches OpenGL.  Harness file, ours. */
