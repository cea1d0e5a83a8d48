/* Test-infrastructure stub, NOT product code.
 * The reference's CPU-side resource classes include <volk/volk.h> only to name a few Vulkan
 * handle/enum types in struct members (src/engine/graphics/GPUObjects.hpp:8). No Vulkan SDK exists
 * in this image, so the oracle build puts this directory on the include path ahead of src/libs.
 * Nothing here is ever called. */
#pragma once
#include <cstdint>
#ifndef VK_NULL_HANDLE
#define VK_NULL_HANDLE nullptr
#endif
#ifndef VK_DEFINE_HANDLE
#define VK_DEFINE_HANDLE(object) typedef struct object##_T* object;
#endif
VK_DEFINE_HANDLE(VkImage)
VK_DEFINE_HANDLE(VkImageView)
VK_DEFINE_HANDLE(VkBuffer)
typedef int VkFormat;
typedef uint32_t VkImageAspectFlags;
typedef int VkImageLayout;
typedef uint64_t VkDeviceSize;
struct VkVertexInputBindingDescription { uint32_t binding, stride; int inputRate; };
struct VkVertexInputAttributeDescription { uint32_t location, binding; int format; uint32_t offset; };
