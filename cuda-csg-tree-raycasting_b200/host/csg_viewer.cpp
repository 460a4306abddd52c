// csg_viewer — OPTIONAL interop viewer: the SDL2/OpenGL window of the reference kept as a thin shell over
// libcsg_b200.  Off by default (`make viewer`; needs SDL2 and an OpenGL 3.0+ driver, neither of which is in the build image —
// the file is syntax-checked against the SDL2 headers the reference vendors, tests/test_host.py, and has not been run here).
//
// What it replaces, with the same behaviour where the reference defines one:
//   Application::CreateAppWindow / Run / Input   (Application.cpp:9-127)        window, frame loop, camera motion per frame
//   InputManager::Input                          (Controls/InputManager.cpp:3-88) W/S A/D Space/LShift, left mouse drag
//   RenderManager ctor / ChangeSize / CalculateRays / RenderRaysData (RenderManager.cpp:3-84, 136-159)
//                                                float4 pixel-unpack buffer registered with CUDA, mapped every frame and
//                                                handed to Raycaster::Raycast, then texture upload + framebuffer blit
// Not carried over: ImGui (the light sliders become the arrow keys, the file dialog becomes argv / drag-and-drop of a scene
// file onto the window); the FPS read-out goes to the window title.
//
//   csg_viewer [scene.txt] [--w 800] [--h 600] [--gpus N]
#include <SDL.h>
#include <SDL_opengl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "csg_raycaster.hpp"

// The four CUDA-GL interop entry points of libcudart (cuda_gl_interop.h), declared here so that the viewer needs neither the
// CUDA headers nor <GL/gl.h> at build time; it links -lcudart.
extern "C" {
struct cudaGraphicsResource;
int cudaGraphicsGLRegisterBuffer(cudaGraphicsResource** resource, unsigned int buffer, unsigned int flags);
int cudaGraphicsUnregisterResource(cudaGraphicsResource* resource);
int cudaGraphicsMapResources(int count, cudaGraphicsResource** resources, void* stream);
int cudaGraphicsUnmapResources(int count, cudaGraphicsResource** resources, void* stream);
int cudaGraphicsResourceGetMappedPointer(void** devPtr, size_t* size, cudaGraphicsResource* resource);
}
static const unsigned int kCudaGraphicsMapFlagsWriteDiscard = 2;   // cudaGraphicsMapFlagsWriteDiscard, RenderManager.cpp:50

using namespace csg_b200;

namespace {

// OpenGL entry points above 1.1 come from the driver (the reference uses glad for the same purpose, RenderManager.cpp:13)
struct GL {
    PFNGLGENBUFFERSPROC GenBuffers;
    PFNGLBINDBUFFERPROC BindBuffer;
    PFNGLBUFFERDATAPROC BufferData;
    PFNGLDELETEBUFFERSPROC DeleteBuffers;
    PFNGLGENFRAMEBUFFERSPROC GenFramebuffers;
    PFNGLBINDFRAMEBUFFERPROC BindFramebuffer;
    PFNGLFRAMEBUFFERTEXTURE2DPROC FramebufferTexture2D;
    PFNGLBLITFRAMEBUFFERPROC BlitFramebuffer;
    PFNGLDELETEFRAMEBUFFERSPROC DeleteFramebuffers;
    bool load()
    {
#define CSG_GL(name) name = reinterpret_cast<decltype(name)>(SDL_GL_GetProcAddress("gl" #name)); if (!name) return false
        CSG_GL(GenBuffers); CSG_GL(BindBuffer); CSG_GL(BufferData); CSG_GL(DeleteBuffers); CSG_GL(GenFramebuffers);
        CSG_GL(BindFramebuffer); CSG_GL(FramebufferTexture2D); CSG_GL(BlitFramebuffer); CSG_GL(DeleteFramebuffers);
#undef CSG_GL
        return true;
    }
};

struct Viewer {
    SDL_Window* window = nullptr;
    SDL_GLContext context = nullptr;
    GL gl{};
    int width = 0, height = 0, gpus = 1;
    GLuint pbo = 0, texture = 0, framebuffer = 0;
    cudaGraphicsResource* pbo_resource = nullptr;
    Raycaster raycaster;
    CSGTree tree;
    bool tree_set = false;
    Camera cam;
    DirectionalLight light;

    // RenderManager::ChangeSize (RenderManager.cpp:136-159): new context for the new size, new PBO + texture, re-registered
    void change_size()
    {
        if (tree_set) raycaster.ChangeSize(width, height, tree, gpus);
        if (pbo_resource) { cudaGraphicsUnregisterResource(pbo_resource); pbo_resource = nullptr; }
        if (texture) glDeleteTextures(1, &texture);
        if (pbo) gl.DeleteBuffers(1, &pbo);
        gl.GenBuffers(1, &pbo);
        gl.BindBuffer(GL_PIXEL_UNPACK_BUFFER, pbo);
        gl.BufferData(GL_PIXEL_UNPACK_BUFFER, (GLsizeiptr)width * height * 4 * (GLsizeiptr)sizeof(float), nullptr, GL_DYNAMIC_DRAW);
        glGenTextures(1, &texture);
        glBindTexture(GL_TEXTURE_2D, texture);
        glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA32F, width, height, 0, GL_RGBA, GL_FLOAT, nullptr);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE);
        gl.BindBuffer(GL_PIXEL_UNPACK_BUFFER, 0);
        if (cudaGraphicsGLRegisterBuffer(&pbo_resource, pbo, kCudaGraphicsMapFlagsWriteDiscard) != 0) {
            std::fprintf(stderr, "cudaGraphicsGLRegisterBuffer failed (is the GL context on an NVIDIA GPU?)\n");
            std::exit(EXIT_FAILURE);
        }
    }

    // Application::LoadCSGTree (Application.cpp:59-83): a scene that does not parse leaves the current one in place
    void load_tree(const std::string& path)
    {
        try {
            std::ifstream in(path.c_str(), std::ios::in);
            if (!in.is_open()) throw std::runtime_error("File not found, or couldn't be open");
            std::stringstream buffer;
            buffer << in.rdbuf();
            CSGTree parsed = CSGTree::Parse(buffer.str());
            tree = std::move(parsed);
            tree_set = true;
            change_size();   // RenderManager::SetTreeToRender (RenderManager.cpp:161-166)
        } catch (const std::exception& exc) {
            std::fprintf(stderr, "Cannot load tree: %s\n", exc.what());
        }
    }

    // RenderManager::CalculateRays + RenderRaysData (RenderManager.cpp:58-84): the mapped PBO is the float4 frame
    void render()
    {
        int w, h;
        SDL_GL_GetDrawableSize(window, &w, &h);
        if (w != width || h != height) { width = w; height = h; change_size(); }
        if (!tree_set || width < 2 || height < 2) { glClearColor(0.08f, 0.08f, 0.11f, 1.0f); glClear(GL_COLOR_BUFFER_BIT); return; }
        void* d_ptr = nullptr;
        cudaGraphicsMapResources(1, &pbo_resource, nullptr);
        cudaGraphicsResourceGetMappedPointer(&d_ptr, nullptr, pbo_resource);
        raycaster.Raycast(d_ptr, cam, light);   // one fused kernel pair; synchronous like the reference's
        cudaGraphicsUnmapResources(1, &pbo_resource, nullptr);

        gl.BindBuffer(GL_PIXEL_UNPACK_BUFFER, pbo);
        glBindTexture(GL_TEXTURE_2D, texture);
        glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, width, height, GL_RGBA, GL_FLOAT, nullptr);
        gl.BindBuffer(GL_PIXEL_UNPACK_BUFFER, 0);
        gl.BindFramebuffer(GL_FRAMEBUFFER, framebuffer);
        gl.FramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, texture, 0);
        gl.BindFramebuffer(GL_READ_FRAMEBUFFER, framebuffer);
        gl.BindFramebuffer(GL_DRAW_FRAMEBUFFER, 0);
        gl.BlitFramebuffer(0, 0, width, height, 0, 0, width, height, GL_COLOR_BUFFER_BIT, GL_NEAREST);   // row 0 = bottom, as rendered
        gl.BindFramebuffer(GL_FRAMEBUFFER, 0);
    }

    void clean_up()
    {  // RenderManager::CleanUp (RenderManager.cpp:169-176)
        raycaster.CleanUp();
        if (pbo_resource) cudaGraphicsUnregisterResource(pbo_resource);
        if (texture) glDeleteTextures(1, &texture);
        if (pbo) gl.DeleteBuffers(1, &pbo);
        if (framebuffer) gl.DeleteFramebuffers(1, &framebuffer);
        if (context) SDL_GL_DeleteContext(context);
        if (window) SDL_DestroyWindow(window);
        SDL_Quit();
    }
};

}  // namespace

int main(int argc, char** argv)
{
    int w = 800, h = 600;   // Application.h:16-17
    Viewer v;
    const char* scene = nullptr;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--w") && i + 1 < argc) w = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--h") && i + 1 < argc) h = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) v.gpus = std::atoi(argv[++i]);
        else scene = argv[i];
    }
    if (SDL_Init(SDL_INIT_VIDEO) < 0) { std::fprintf(stderr, "SDL cannot initialize video subsytem\n"); return EXIT_FAILURE; }
    SDL_GL_SetAttribute(SDL_GL_CONTEXT_MAJOR_VERSION, 4);   // Application.cpp:17-22
    SDL_GL_SetAttribute(SDL_GL_CONTEXT_MINOR_VERSION, 1);
    SDL_GL_SetAttribute(SDL_GL_CONTEXT_PROFILE_MASK, SDL_GL_CONTEXT_PROFILE_CORE);
    SDL_GL_SetAttribute(SDL_GL_DOUBLEBUFFER, 1);
    v.window = SDL_CreateWindow("CSG RayCasting (libcsg_b200)", 100, 100, w, h, SDL_WINDOW_OPENGL | SDL_WINDOW_RESIZABLE);
    if (!v.window) { std::fprintf(stderr, "SDL_CreateWindow: %s\n", SDL_GetError()); return EXIT_FAILURE; }
    v.context = SDL_GL_CreateContext(v.window);
    if (!v.context || !v.gl.load()) { std::fprintf(stderr, "Cannot create an OpenGL 4.1 core context\n"); return EXIT_FAILURE; }
    SDL_GL_SetSwapInterval(1);
    SDL_EventState(SDL_DROPFILE, SDL_ENABLE);
    v.gl.GenFramebuffers(1, &v.framebuffer);
    SDL_GL_GetDrawableSize(v.window, &v.width, &v.height);
    v.change_size();
    if (scene) v.load_tree(scene);

    bool quit = false, dragging = false;
    int forward = 0, backward = 0, left = 0, right = 0, up = 0, down = 0;   // InputManager::camControls
    Uint32 old_time = SDL_GetTicks(), frames = 0, fps_t0 = old_time;
    while (!quit) {
        int rel_x = 0, rel_y = 0;
        SDL_Event e;
        while (SDL_PollEvent(&e) != 0) {   // InputManager::Input (InputManager.cpp:3-88)
            if (e.type == SDL_QUIT) quit = true;
            if (e.type == SDL_MOUSEMOTION) { rel_x = e.motion.xrel; rel_y = e.motion.yrel; }
            if (e.type == SDL_MOUSEBUTTONDOWN && e.button.button == SDL_BUTTON_LEFT) { dragging = true; SDL_SetRelativeMouseMode(SDL_TRUE); }
            if (e.type == SDL_MOUSEBUTTONUP && e.button.button == SDL_BUTTON_LEFT) { dragging = false; SDL_SetRelativeMouseMode(SDL_FALSE); }
            if (e.type == SDL_DROPFILE) { v.load_tree(e.drop.file); SDL_free(e.drop.file); }
            if ((e.type == SDL_KEYDOWN || e.type == SDL_KEYUP) && e.key.repeat == 0) {
                const int pressed = e.key.state == SDL_PRESSED;
                switch (e.key.keysym.sym) {
                    case SDLK_w: forward = pressed; break;
                    case SDLK_s: backward = pressed; break;
                    case SDLK_a: left = pressed; break;
                    case SDLK_d: right = pressed; break;
                    case SDLK_SPACE: up = pressed; break;
                    case SDLK_LSHIFT: down = pressed; break;
                    case SDLK_ESCAPE: quit = true; break;
                    default: break;
                }
            }
            if (e.type == SDL_KEYDOWN) {   // the ImGui sliders of the reference (RenderManager.cpp:103-105): 2 degrees per key press
                const float step = 2.0f * 3.14159f / 180.0f;
                switch (e.key.keysym.sym) {
                    case SDLK_UP: v.light.polar += step; break;
                    case SDLK_DOWN: v.light.polar -= step; break;
                    case SDLK_LEFT: v.light.azimuth -= step; break;
                    case SDLK_RIGHT: v.light.azimuth += step; break;
                    default: break;
                }
            }
        }
        // Application::Input (Application.cpp:85-127): 0.1 units per frame and key, 0.005 rad per pixel of mouse motion
        const float move_forward = 0.1f * (forward - backward), move_right = 0.1f * (right - left), move_up = 0.1f * (up - down);
        if (dragging) v.cam.rotate(-0.005f * rel_y, -0.005f * rel_x);
        v.cam.move(move_forward, move_right, move_up);

        try {
            v.render();
        } catch (const std::exception& exc) {
            std::fprintf(stderr, "render: %s\n", exc.what());
            break;
        }
        SDL_GL_SwapWindow(v.window);
        ++frames;
        const Uint32 now = SDL_GetTicks();
        if (now - fps_t0 >= 500) {   // the reference averages its FPS read-out over a cyclic buffer (Application.cpp:44-50)
            char title[128];
            std::snprintf(title, sizeof title, "CSG RayCasting (libcsg_b200) — %.1f FPS, %dx%d", 1000.0f * frames / (now - fps_t0), v.width, v.height);
            SDL_SetWindowTitle(v.window, title);
            frames = 0;
            fps_t0 = now;
        }
        old_time = now;
    }
    (void)old_time;
    v.clean_up();
    return 0;
}
