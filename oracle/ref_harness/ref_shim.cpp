/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Link-time stand-ins for the Vulkan-facing members that the reference's CPU-side loaders reference but that the
 * offline path tracer never executes with Storage::CPU (src/engine/resources/ResourcesManager.cpp, Mesh.cpp,
 * Scene.cpp call them only behind `options & Storage::GPU`). Every body is a no-op or the trivial CPU part of the
 * original; none of them is on the path being checked. The one exception, Texture::sampleCubemap, IS on the path
 * (PathTracer.cpp:35,49): its body is not retyped here but included from a file sliced out of the reference tree
 * at build time (oracle/build_ref.sh -> oracle/_ref/gen/sampleCubemap.inc).
 */
#include "resources/Texture.hpp"
#include "resources/Buffer.hpp"
#include "resources/Font.hpp"
#include "resources/Mesh.hpp"
#include "graphics/GPUObjects.hpp"
#include "graphics/GPU.hpp"
#include "graphics/Program.hpp"
#include "renderers/DebugViewer.hpp"
#include "system/System.hpp"

Texture::Texture(const std::string & name) : _name(name) {}
Texture::Texture(Texture &&) = default;
Texture::~Texture() { images.clear(); }
void Texture::upload(const Layout &, bool) {}
void Texture::clearImages() { images.clear(); }
void Texture::clean() { images.clear(); }
const std::string & Texture::name() const { return _name; }

#include "sampleCubemap.inc"

/* Every scene texture is decoded to four channels on the CPU (the survey checked this against
 * VkUtils::getGPULayout for the layouts Codable::decodeTexture produces). */
unsigned int GPUTexture::getChannelsCount(const Layout &) { return 4; }
void GPUMesh::clean() {}
void GPU::setupMesh(Mesh &) {}
void DebugViewer::trackDefault(const Mesh *) {}
void DebugViewer::untrackDefault(const Mesh *) {}

Buffer::Buffer(size_t sizeInBytes, BufferType atype, const std::string & name) : type(atype), _name(name) {}
void Buffer::upload(size_t, unsigned char *, size_t) {}
Buffer::~Buffer() {}

Program::Program(const std::string & name, const std::string &, const std::string &, const std::string &, const std::string &) : _name(name), _type(Type::GRAPHICS) {}
Program::Program(const std::string & name, const std::string &) : _name(name), _type(Type::COMPUTE) {}
void Program::reload(const std::string &, const std::string &, const std::string &, const std::string &) {}
void Program::reload(const std::string &) {}
void Program::clean() {}

Font::Font(std::istream &) {}

const char * System::widen(const std::string & str) { return str.c_str(); }
std::string System::narrow(char * str) { return std::string(str); }
