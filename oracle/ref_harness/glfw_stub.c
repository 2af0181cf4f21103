/* No-op definitions of the 24 GLFW entry points the reference's windowed branch links against
 * (reference: src/util/Window.cpp, src/main.cpp:231-253,298-437).  The headless branch (no -window)
 * never calls them; this only lets the UNMODIFIED reference sources link into oracle/_ref/ref_pt on a
 * box with no GLFW/OpenGL.  Harness file, ours; C symbols carry no signature so one shape serves all. */
#define STUB(name) void *name(void) { return 0; }
STUB(glfwCreateWindow) STUB(glfwDestroyWindow) STUB(glfwGetFramebufferSize) STUB(glfwGetProcAddress)
STUB(glfwGetWindowUserPointer) STUB(glfwInit) STUB(glfwMakeContextCurrent) STUB(glfwPollEvents)
STUB(glfwSetCharCallback) STUB(glfwSetCursorEnterCallback) STUB(glfwSetCursorPosCallback)
STUB(glfwSetFramebufferSizeCallback) STUB(glfwSetInputMode) STUB(glfwSetJoystickCallback)
STUB(glfwSetKeyCallback) STUB(glfwSetMouseButtonCallback) STUB(glfwSetScrollCallback)
STUB(glfwSetWindowTitle) STUB(glfwSetWindowUserPointer) STUB(glfwSwapBuffers) STUB(glfwTerminate)
STUB(glfwWindowHint) STUB(glfwWindowShouldClose)
double glfwGetTime(void) { return 0.0; }
